"""The HDF5 writer moved into the package (nanoreviser_b200/h5write.py: the training path saves Keras-layout weight files with it);
the fixture generators and tests keep importing it from here."""
from nanoreviser_b200.h5write import *  # noqa: F401,F403
from nanoreviser_b200.h5write import Writer, write_tree  # noqa: F401
