cd /root/repo
for env in "A=1" "NR_OWN_GEMM=0" "NR_EARLY_FIFO=0" "NR_TC2_HALVES=1"; do
  echo "== $env"
  env $env timeout 300 python -m pytest "tests/test_gpu_install.py::test_trainer_loop_through_the_rebound_forward" -q -p no:cacheprovider -rP -k "bf16 and not x3" 2>&1 | grep -E "^step|passed|failed" 
done
