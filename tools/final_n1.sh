set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_reference_arm.json 2> gpurun_out/r02c_ref.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r02c_bench_B256.json 2> gpurun_out/r02c_bench_B256.err
python bench.py --workload cfg5 --steps 10 --warmup 3 > gpurun_out/r02c_bench_cfg5_B148.json 2> gpurun_out/r02c_bench_cfg5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02c_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pso_sliced --launch-skip 1 -c 1 -o gpurun_out/r02c_pso_sliced -f python tools/prof_run.py 256 2 > gpurun_out/r02c_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pso_sliced --launch-skip 1 -c 1 -o gpurun_out/r02c_pso_sliced_cfg5 -f python tools/prof_cfg5.py 2 > gpurun_out/r02c_ncu_full_cfg5.log 2>&1
python tools/phase_split.py 256 > gpurun_out/r02c_phase_split.txt 2>&1
python tools/e2e_depth.py 256 40 > gpurun_out/r02c_e2e_depth.txt 2>&1
python tools/e2e_split.py > gpurun_out/r02c_e2e_split.txt 2>&1
tail -c 400 gpurun_out/r02c_bench_B256.err
python -c "
import json
for f in ['r02c_bench_B256','r02c_bench_cfg5_B148','r02c_bench_reference_arm']:
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d.get('value'), d.get('e2e'), d.get('parity',{}).get('ok'), d.get('clocks'))
"
