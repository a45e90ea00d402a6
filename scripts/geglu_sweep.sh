for shape in "65536 2560 320" "16384 5120 640" "4096 10240 1280" "16384 2560 320" "4096 5120 640" "1024 10240 1280" "8192 2560 320"; do
  for bn in 128 160 256; do
    timeout 60 python scripts/gemm_micro.py $shape 1 $bn 3 0 2>&1 | tail -1
  done
done
