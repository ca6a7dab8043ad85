python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reference_library_itself" 2>&1 | tail -3
export QLB_LIBRARY=$PWD/qkd_ldpc_b200/lib/variants/liblibm.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_codes.py -m gpu -x -q -k "(campaign and f64) or golden or waterfall_fp64 or no_clamp or edge_cases or small_codes or kat or streaming_fp64 or reference_library_itself" > gpurun_out/r2x_libm_forms.log 2>&1; tail -5 gpurun_out/r2x_libm_forms.log
python scripts/profile_point.py f64 0.09 2960 | tail -1 >> gpurun_out/r2x_libm_forms.log; tail -1 gpurun_out/r2x_libm_forms.log
