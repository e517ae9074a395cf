set -x
python tools/time_stream.py T10 30 > gpurun_out/r02h_stream.txt 2>&1
cat gpurun_out/r02h_stream.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02h_gputests.txt
cat gpurun_out/r02h_gputests.txt
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 50 python tools/sanitize_run.py > gpurun_out/r02h_racecheck2.txt 2>&1
tail -3 gpurun_out/r02h_racecheck2.txt
