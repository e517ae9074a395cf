python -m pytest tests/test_loader_gpu.py tests/test_native_abi.py -m "gpu or not gpu" -x -q -k "read_paths or host_copy or batch_front or load_files" 2>&1 | tail -3
for mode in preadv mmap; do for t in 3 8; do echo "mode $mode threads $t"; MS_B200_READ_MODE=$mode MS_B200_READ_THREADS=$t python tools/time_pipeline_files.py T127 8 2>&1 | grep "load_files only\|overlapped" | tail -2; done; done
