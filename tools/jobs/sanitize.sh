for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_driver.py 2>&1 | grep -v "^$" | tail -8
done
