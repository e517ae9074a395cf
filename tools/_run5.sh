set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02m_gputests.txt
cat gpurun_out/r02m_gputests.txt
python tools/time_pipeline_files.py T127 8 2>&1 | grep "ms/trial" > gpurun_out/r02m_pipe.txt
cat gpurun_out/r02m_pipe.txt
