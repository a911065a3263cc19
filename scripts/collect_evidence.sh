#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun): bench lines of every workload, the reference
# arm, ncu launch lists of one drop-in step, one `ncu --set full` capture of the dominant kernels.
#   gpurun -- 'bash scripts/collect_evidence.sh r2b'     ->  gpurun_out/<tag>_*
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
for W in voc_scribble_b1 voc_scribble_b4 voc_tag_b2 densepose_b1 voc_scribble_softmax_b1; do
  timeout 600 python bench.py --no-train-arm --workload $W > $O/${TAG}_bench_$W.json 2> $O/${TAG}_bench_$W.err
done
timeout 900 python bench.py --impl reference > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
timeout 900 python bench.py --workload train_voc_b4 > $O/${TAG}_train_voc_b4_1gpu.json 2> $O/${TAG}_train_voc_b4_1gpu.err
for W in voc_scribble_b1 voc_scribble_b4; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/${TAG}_launches_$W.csv python scripts/profile_step.py --workload $W --steps 3 > /dev/null 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kmeans_small -s 2 -c 1 -f \
  -o $O/${TAG}_full_kmeans_b1 python scripts/profile_step.py --workload voc_scribble_b1 --steps 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_normalize_pack -s 2 -c 1 -f \
  -o $O/${TAG}_full_prepass_b1 python scripts/profile_step.py --workload voc_scribble_b1 --steps 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kmeans_small -s 2 -c 1 -f \
  -o $O/${TAG}_full_kmeans_b4 python scripts/profile_step.py --workload voc_scribble_b4 --steps 3 > /dev/null 2>&1
ls -la $O | tail -30
for W in voc_scribble_b1 voc_scribble_b4 voc_tag_b2 densepose_b1 voc_scribble_softmax_b1; do
  python - <<PY
import json
d=json.loads(open('$O/${TAG}_bench_$W.json').read().strip().splitlines()[-1])
print('$W', 'resident %.3f ms, e2e %.3f ms, reference %.1f ms, kmeans %.1f us' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['cpu_baseline']['ms_per_step'], 1e3*d['roofline']['all'].get('spml_kmeans',{}).get('ms_per_call',0)))
PY
done
tail -c 400 $O/${TAG}_bench_reference.json; echo; tail -c 600 $O/${TAG}_train_voc_b4_1gpu.json
