set -x
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > gpurun_out/r02p_smoke.txt 2>&1
tail -2 gpurun_out/r02p_smoke.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02p_gputests.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02p_bench_ref.json 2> gpurun_out/r02p_bench_ref.err
python bench.py > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02p_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02p_b.log 2>&1
python tools/step_timeline.py > gpurun_out/r02p_timeline.txt 2>&1
cat gpurun_out/r02p_gputests.txt; tail -c 600 gpurun_out/r02p_bench.json
