#!/bin/bash
# precision experiment: rebuild nrv_fused_pair.cu with -DNRV_FP_SKIP_XLO=<mask> and report max |dP| / label agreement on the unitest set
for m in "$@"; do
  echo "=== NRV_FP_SKIP_XLO=$m"
  touch nanoreviser_b200/csrc/nrv_fused_pair.cu
  if [ "$m" = "0" ]; then python -m nanoreviser_b200.build > /dev/null; else NRV_EXTRA_NVCC=-DNRV_FP_SKIP_XLO=$m python -m nanoreviser_b200.build > /dev/null; fi
  python - <<'P'
import numpy as np, os, glob
from nanoreviser_b200 import api, engine, fast5, weights
reads=[fast5.read_fast5_arrays(f) for f in sorted(glob.glob('tests/golden/fast5/*.fast5'))]
for sp in ("ecoli","human"):
    m1,m2=weights.load_species(sp,'model')
    gold=np.load('tests/golden/forward_%s.npz'%sp)
    with engine.Reviser(m1,m2,device=0) as rv:
        out=api.revise_reads(reads,reviser=rv,want_labels=True,want_probs=True)
    w0=0; d1=d2=0; flips=0; tot=0; same_seq=0
    for k,r in enumerate(reads):
        M=r.n_bases-11
        d1=max(d1,np.abs(out.p1[w0:w0+M]-gold['r%d_P1_f64'%k]).max()); d2=max(d2,np.abs(out.p2[w0:w0+M]-gold['r%d_P2_f64'%k]).max())
        flips+=int((out.y1[w0:w0+M]!=gold['r%d_y1_f64'%k]).sum()+(out.y2[w0:w0+M]!=gold['r%d_y2_f64'%k]).sum()); tot+=2*M
        same_seq+=int(out.sequence(k)==gold['r%d_revised'%k].tobytes().decode())
        w0+=M
    print("%s: max|dP1| %.2e max|dP2| %.2e label flips %d / %d, identical sequences %d/5" % (sp,d1,d2,flips,tot,same_seq))
P
done
touch nanoreviser_b200/csrc/nrv_fused_pair.cu; python -m nanoreviser_b200.build > /dev/null
