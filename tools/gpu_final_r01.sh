mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['whole_models']['picnet_ref_256']['value'], d['whole_models']['refpsp_1024']['value'], d['clocks'])"
python tools/debug/one_picnet.py 8 fp32 > gpurun_out/one_picnet.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_picnet.csv python tools/debug/one_picnet.py 8 fp32 > gpurun_out/ncu_picnet.log 2>&1
wc -l gpurun_out/launches_picnet.csv
timeout 600 python tools/perf/perf_picnet.py > gpurun_out/perf_picnet.txt 2>&1; tail -4 gpurun_out/perf_picnet.txt
timeout 300 python tools/perf/perf_picnet_breakdown.py 2>&1 | head -2 | tee gpurun_out/perf_picnet_pieces.txt
timeout 400 python tools/perf/perf_graphs.py > gpurun_out/perf_graphs.txt 2> gpurun_out/perf_graphs.err; tail -3 gpurun_out/perf_graphs.txt
