"""TEST INFRASTRUCTURE -- import-only stub so reference modules that `import matplotlib...`
at module top (mpc.py:2, human_avoidance_env.py:6-9, HA_and_PT env :7-14) can be imported.
Nothing on the hot path draws."""


class _Anything:
    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def __getattr__(name):
    return _Anything()
