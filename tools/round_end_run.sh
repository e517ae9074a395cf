# What is run on a B200 at the end of a round (through gpurun): smoke, GPU tests, both bench arms, the ncu launch list of the bench
# command, the step timeline.  Outputs go to gpurun_out/rXX_* (rename per round) and are summarised into profiles/.
set -x
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > gpurun_out/rXX_smoke.txt 2>&1
tail -2 gpurun_out/rXX_smoke.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/rXX_gputests.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/rXX_bench_ref.json 2> gpurun_out/rXX_bench_ref.err
python bench.py > gpurun_out/rXX_bench.json 2> gpurun_out/rXX_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/rXX_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/rXX_b.log 2>&1
python tools/step_timeline.py > gpurun_out/rXX_timeline.txt 2>&1
cat gpurun_out/rXX_gputests.txt; tail -c 600 gpurun_out/rXX_bench.json
