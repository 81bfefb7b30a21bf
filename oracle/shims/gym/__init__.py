"""TEST INFRASTRUCTURE -- import-only stub: reference crowd_sim.py:4,21 only subclasses gym.Env."""


class Env:
    pass
