(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2n_tests.log 2>&1; echo "ALL gpu tests rc=$?"; tail -6 gpurun_out/r2n_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
