# compile-time sweep of field_tc_kernel variants on the GPU box: bash tools/sweep_field_tc.sh "-DPNERF_EPW=4" "-DPNERF_F32X2=0" ...
for defs in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --fmad=true $defs -c pointnerf2studio_b200/csrc/field_tc.cu -o pointnerf2studio_b200/build/field_tc.o && \
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o pointnerf2studio_b200/libpnerf_b200.so pointnerf2studio_b200/build/*.o && \
  echo "== $defs" && timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"field": [0-9.]*'
done
