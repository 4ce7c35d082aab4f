# usage: bash tools/trip_tests.sh <N> [pytest -k expr]
N=${1:-2}
mkdir -p gpurun_out
export PDT_TEST_WORLD=$N
export PDT_TEST_EXPERIMENTAL=${PDT_TEST_EXPERIMENTAL:-1}   # include the tests that have not run on hardware yet
timeout -s KILL 900 python -m pytest tests/test_gpu_multigpu.py -q -m gpu --timeout 240 -p no:cacheprovider ${2:+-k "$2"} > gpurun_out/comm_tests_$N.log 2>&1
grep -n "^E  \|passed\|failed" gpurun_out/comm_tests_$N.log | cut -c1-400 | tail -n 40
