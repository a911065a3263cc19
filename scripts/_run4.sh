export SPML_B200_BINDING=ctypes
for rnd in 1 2; do
for L in libspml_b200.so libspml_b200_t64.so libspml_b200_t32.so; do
  echo -n "$L  "; SPML_B200_LIB=spml_b200/$L python scripts/ab_step.py --child voc_scribble_b1 2>&1 | tail -1
done; done
for L in libspml_b200.so libspml_b200_t64.so; do
  echo -n "b4 $L  "; SPML_B200_LIB=spml_b200/$L python scripts/ab_step.py --child voc_scribble_b4 2>&1 | tail -1
done
SPML_B200_LIB=spml_b200/libspml_b200_t64.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -x -q 2>&1 | tail -2
