set -u
for tool in memcheck initcheck racecheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_k1_k3.py k2"
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_k1_k3.py k2 2>&1 | grep -v "^$" | tail -12
done > gpurun_out/sanitizer_k2.txt 2>&1
tail -40 gpurun_out/sanitizer_k2.txt | cut -c1-300
bash tools/gpu_profile.sh etc1 > /dev/null 2>&1; tail -2 gpurun_out/prof_etc1.log
