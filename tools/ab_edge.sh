timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_expand_state.py tests/test_gpu_validity.py -x -q -s 2>&1 | grep -E "passed|failed|LazyARAStar|unchanged reference|Error|assert"
