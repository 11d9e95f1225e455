# where does CUDA context creation go? (diagnostic) -- five fresh processes, hsgpu_ctx_create only
for i in 1 2 3 4 5; do
  HSGPU_TIMING=1 python - <<'PY' 2>&1 | grep "hsgpu timing\|total"
import time, ctypes as C, os, sys
t0 = time.perf_counter()
L = C.CDLL(os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "hairsplitter_b200", "libhsgpu.so"))
t1 = time.perf_counter()
h = C.c_void_p()
L.hsgpu_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
rc = L.hsgpu_ctx_create(0, C.byref(h))
t2 = time.perf_counter()
print(f"total: dlopen {1e3*(t1-t0):.1f} ms, hsgpu_ctx_create {1e3*(t2-t1):.1f} ms rc={rc}")
os._exit(0)
PY
done
