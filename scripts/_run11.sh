timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_ring.log 2>&1
grep -v "^=========     Host Frame\|^=========         in \|libcuda\|libtorch\|python" gpurun_out/san_ring.log | head -60
