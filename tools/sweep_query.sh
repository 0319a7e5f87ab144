# compile-time sweep of query_kernel on the GPU box: bash tools/sweep_query.sh "-DPNERF_Q_MINB=10" ...
for defs in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --fmad=true $defs -c pointnerf2studio_b200/csrc/query.cu -o pointnerf2studio_b200/build/query.o 2>/dev/null && \
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o pointnerf2studio_b200/libpnerf_b200.so pointnerf2studio_b200/build/*.o && \
  echo "== $defs" && timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train --no-stress 2>&1 | tail -1 | grep -o '"query": [0-9.]*'
done
