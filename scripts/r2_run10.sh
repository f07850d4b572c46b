(time timeout 600 python -m pytest tests/test_gpu_cli_verbose.py -m gpu -x -q) > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2i_tests.log
stabilizer_gs_b200/bin/klocal_sparse -v1 examples/hamiltonian_finite.txt 2>/dev/null | sed -n 1,60p
