for tool in memcheck initcheck racecheck; do
  echo "== $tool"
  timeout 500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_k1_k3.py > gpurun_out/sanitizer_$tool.txt 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok" gpurun_out/sanitizer_$tool.txt | tail -8
done
