mkdir -p gpurun_out
export SPML_B200_KMEANS=cluster
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -q -x -k "kmeans_matches_oracle or kmeans_with_empty or fixed_point" > gpurun_out/san_mem_cluster.txt 2>&1
tail -4 gpurun_out/san_mem_cluster.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_ops.py -q -x -k "kmeans_matches_oracle or kmeans_with_empty" > gpurun_out/san_race_cluster.txt 2>&1
tail -4 gpurun_out/san_race_cluster.txt
unset SPML_B200_KMEANS
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k golden > gpurun_out/san_mem_parity.txt 2>&1
tail -4 gpurun_out/san_mem_parity.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "golden and not softmax and not densepose" > gpurun_out/san_race_parity.txt 2>&1
tail -4 gpurun_out/san_race_parity.txt
