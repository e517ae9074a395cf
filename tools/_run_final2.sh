set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02m_gputests.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02m_bench_ref.json 2> gpurun_out/r02m_bench_ref.err
python bench.py > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02m_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02m_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ms_load_kernel -s 2 -c 1 -o gpurun_out/r02m_fused -f python tools/profile_run.py T10 4 > gpurun_out/r02m_ncu.log 2>&1
python tools/step_timeline.py > gpurun_out/r02m_timeline.txt 2>&1
cat gpurun_out/r02m_gputests.txt; cat gpurun_out/r02m_timeline.txt; tail -c 1500 gpurun_out/r02m_bench.json
