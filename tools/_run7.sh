timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | tail -3
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_run.py 2>&1 | tail -2
python tools/fuzz_loader.py 2>&1 | tail -3
