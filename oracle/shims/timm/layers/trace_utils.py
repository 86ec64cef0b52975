def _assert(condition, message):
    assert condition, message
