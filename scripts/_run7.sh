export SPML_B200_BINDING=ctypes
for rnd in 1 2 3; do
for L in libspml_b200_prev.so libspml_b200.so; do
  echo -n "$L  "; SPML_B200_LIB=spml_b200/$L python scripts/ab_step.py --child voc_scribble_b1 2>&1 | tail -1
done; done
unset SPML_B200_BINDING
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py tests/test_gpu_reference_dropin.py -x -q 2>&1 | tail -2
