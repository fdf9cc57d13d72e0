python scripts/trace_gemm_tc.py
rm -f hulc_b200/lib/libhulc_trace.so
