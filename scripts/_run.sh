bash scripts/bounds_check_build.sh run > gpurun_out/r2n_bounds.log 2>&1; echo "bounds rc=$?"; tail -25 gpurun_out/r2n_bounds.log
python scripts/sanitize_case.py > gpurun_out/r2n_product_agreement.log 2>&1; echo "agreement rc=$?"; tail -3 gpurun_out/r2n_product_agreement.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_gpu_tests.log 2>&1; tail -5 gpurun_out/r2n_gpu_tests.log
