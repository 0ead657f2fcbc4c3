set -x
mkdir -p gpurun_out/r2f
cp hipstr_b200/libhipstr_b200.so /tmp/lib_keep.so
for v in nosplit split splittr splittrah splitah nosplit; do
  cp tools/_variants/libhipstr_b200_$v.so hipstr_b200/libhipstr_b200.so
  echo "== $v" >> gpurun_out/r2f/variants.log
  python tools/quick_time.py 200 8 5 2>&1 | grep "run 2\|run 3\|run 4\|checksum" >> gpurun_out/r2f/variants.log
done
cp /tmp/lib_keep.so hipstr_b200/libhipstr_b200.so
cat gpurun_out/r2f/variants.log
