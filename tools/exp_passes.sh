for p in 2 1; do
  NRV_EXTRA_NVCC="-DNRV_REC_PASSES=$p" python -m nanoreviser_b200.build --force > /dev/null 2>&1
  echo "=== REC_PASSES=$p"
  python tests/diag_paths.py tc 2>&1 | grep -v "identical: True"
done
