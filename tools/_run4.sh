set -x
for t in 8 16; do
MS_B200_READ_THREADS=$t python tools/time_pipeline_files.py T127 8 2>&1 | grep "ms/trial" > gpurun_out/r02j_files_$t.txt
cat gpurun_out/r02j_files_$t.txt
done
