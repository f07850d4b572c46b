(time timeout 600 python -m pytest tests/test_gpu_periodic_classes.py tests/test_gpu_python_api.py tests/test_gpu_klocal.py -m gpu -x -q) > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2g_tests.log
(time python py/klocal_periodic.py examples/hamiltonian_periodic.txt) 2>&1 | tail -12
