"""Import shim for the *unmodified* reference env (jesselasse/MARL-DMFB).

Only used by ``tests/golden/make_golden.py`` in the build container, where the
reference is mounted read-only at ``/root/reference``.  Nothing that runs on the
GPU box imports this file: the GPU box has no ``/root/reference``; it only sees
the ``*.npz`` fixtures this shim helped to generate.

The reference imports ``gym``, ``pettingzoo`` and ``numpy.lib.function_base``
(env/DMFB/dmfb.py:11-17, env/MEDA/meda.py:9-12) which are not installed here,
and calls ``random.seed(datetime.now())`` (dmfb.py:154, meda.py:155) which is a
TypeError on Python >= 3.11.  We stub the former as attribute holders and wrap
the latter; no reference source line is changed.
"""
import datetime
import random
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Space:
    def __init__(self, *a, **k):
        self.args, self.kwargs = a, k


def install():
    """Install the stub modules and return (dmfb_module, meda_module)."""
    if "env.DMFB.dmfb" in sys.modules and "env.MEDA.meda" in sys.modules:
        return sys.modules["env.DMFB.dmfb"], sys.modules["env.MEDA.meda"]
    spaces = _mod("gym.spaces", Discrete=_Space, Box=_Space)
    wrappers = _mod("gym.wrappers")
    seeding = _mod("gym.utils.seeding")
    utils = _mod("gym.utils", seeding=seeding)
    error = _mod("gym.error")
    _mod("gym", spaces=spaces, wrappers=wrappers, utils=utils, error=error)

    class ParallelEnv:  # noqa: D401 - empty base class, as in pettingzoo's ABC
        pass

    penv = _mod("pettingzoo.utils.env", ParallelEnv=ParallelEnv)
    putils = _mod("pettingzoo.utils", env=penv)
    _mod("pettingzoo", utils=putils)
    if "numpy.lib.function_base" not in sys.modules:
        _mod("numpy.lib.function_base", select=np.select)

    _orig_seed = random.seed

    def _seed(a=None, version=2):
        if isinstance(a, datetime.datetime):
            a = a.timestamp()
        return _orig_seed(a, version)

    random.seed = _seed
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import env.DMFB.dmfb as ref_dmfb  # noqa: E402
    import env.MEDA.meda as ref_meda  # noqa: E402
    return ref_dmfb, ref_meda


class DrawInjector:
    """Replaces an env module's ``random`` global so that the move-success
    draw of droplet ``i`` at the current step is ``table[i]`` (dmfb.py:335,
    meda.py:280).  Everything else is forwarded to the real ``random`` module.
    The droplet index is published by wrapping ``moveOneDroplet`` on the
    routing-manager *instance* (the class is untouched)."""

    def __init__(self, module, routing_manager):
        self._module = module
        self._real = random
        self._rm = routing_manager
        self.table = None
        self.cur = None
        self.consumed = None
        orig = routing_manager.moveOneDroplet
        inj = self

        def wrapped(droplet_index, *a, **k):
            inj.cur = droplet_index
            try:
                return orig(droplet_index, *a, **k)
            finally:
                inj.cur = None

        routing_manager.moveOneDroplet = wrapped

    def random(self):
        assert self.cur is not None and self.table is not None
        self.consumed[self.cur] += 1
        return float(self.table[self.cur])

    def __getattr__(self, name):
        return getattr(self._real, name)

    def step(self, env, actions, table):
        """env.step(actions) with this step's draws taken from table[A]."""
        self.table = table
        self.consumed = np.zeros(len(table), dtype=np.int64)
        self._module.random = self
        try:
            return env.step(actions)
        finally:
            self._module.random = self._real
            self.table = None
