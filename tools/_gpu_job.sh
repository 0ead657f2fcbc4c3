set -x
mkdir -p gpurun_out/r2d
cp hipstr_b200/libhipstr_b200.so /tmp/lib_keep.so
for v in base s18 s20 i4 kl1 base; do
  cp tools/_variants/libhipstr_b200_$v.so hipstr_b200/libhipstr_b200.so
  echo "== $v" >> gpurun_out/r2d/variants3.log
  python tools/quick_time.py 200 8 4 2>&1 | grep "run 3\|checksum" >> gpurun_out/r2d/variants3.log
done
cp /tmp/lib_keep.so hipstr_b200/libhipstr_b200.so
cat gpurun_out/r2d/variants3.log
