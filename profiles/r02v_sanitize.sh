OUT=gpurun_out/r02v_sanitizer.txt; : > $OUT
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool" >> $OUT
  timeout 500 compute-sanitizer --tool $tool python profiles/sanitize.py 2>&1 | grep -vE "^=+$" | tail -25 >> $OUT
done
tail -60 $OUT
