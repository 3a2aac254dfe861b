mkdir -p gpurun_out
run() {
  echo "=== compute-sanitizer --tool $1 python tools/sanitize_small.py $2"
  timeout $3 compute-sanitizer --tool $1 python tools/sanitize_small.py $2 2>&1 | grep -v "^$" | tail -9
}
{
  run memcheck anysize 200
  run memcheck tiled 300
  run memcheck resident 200
  run racecheck anysize 200
  run racecheck resident 300
  run racecheck tiled 400
} > gpurun_out/r2_compute_sanitizer_v2.txt 2>&1
grep -E "===|ERROR SUMMARY|RACECHECK SUMMARY|done|hazard" gpurun_out/r2_compute_sanitizer_v2.txt | head -40
