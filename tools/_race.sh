set -x
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 200 python tools/sanitize_run.py > gpurun_out/r02h_racecheck.txt 2>&1
grep -c "Race reported\|hazard" gpurun_out/r02h_racecheck.txt
grep -A6 "=========  *\(Error\|Warning\|Race\)" gpurun_out/r02h_racecheck.txt | grep -o "in .*\|at .*" | sort | uniq -c | sort -rn | head -40
tail -5 gpurun_out/r02h_racecheck.txt
