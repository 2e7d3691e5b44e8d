"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Import shim for the *live* Python reference: ``/root/reference`` in the development container, and on the
GPU box (where that path is absent) the byte-compiled, sourceless copy ``oracle/_ref`` that the committed recipe
``oracle/fetch_ref.py`` builds from those sources.  It is used by

  * ``oracle/gen_golden.py``   -- to generate the committed fixtures under ``tests/golden/``;
  * ``tests/test_oracle_vs_reference.py`` -- differential test of the C restatement
    (``oracle/ah_oracle.c``) against the reference, skipped when the reference is absent;
  * ``oracle/time_reference.py`` / ``bench.py`` -- to time the reference's own CPU loop on the host cores.

The reference cannot be imported as shipped in this image: it pulls in ``redis``
(environment/airhockey.py:7, lib/utils/connect.py:8), ``keras`` (lib/utils/noisy_dense.py:4-7 via
lib/utils/__init__.py:2), ``pytz`` (lib/rewards.py:7), ``tensorflow`` (lib/utils/helpers.py:6),
``imp`` (lib/agents/*.py, removed in Python 3.12) and ``np.Infinity`` (lib/rewards.py:31, removed
in NumPy 2).  None of those touch the arithmetic of the environment step, so they are stubbed.
"""
from __future__ import annotations

import datetime
import logging
import os
import sys
import tempfile
import types
from unittest.mock import MagicMock

import numpy as np

_COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


_COMPILED_EXT = ".refc"      # a sourceless .pyc by content (oracle/fetch_ref.py)


def _has_reference(root: str) -> bool:
    return any(os.path.isfile(os.path.join(root, "environment", "airhockey" + ext)) for ext in (".py", _COMPILED_EXT))


class _CompiledFinder:
    """Import finder for the byte-compiled reference: ``pkg/mod.refc`` / ``pkg/__init__.refc`` under ``root``."""

    def __init__(self, root: str):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        import importlib.util
        base = os.path.join(self.root, *fullname.split("."))
        if os.path.isfile(os.path.join(base, "__init__" + _COMPILED_EXT)):
            file = os.path.join(base, "__init__" + _COMPILED_EXT)
            loader = importlib.machinery.SourcelessFileLoader(fullname, file)
            return importlib.util.spec_from_file_location(fullname, file, loader=loader, submodule_search_locations=[base])
        if os.path.isfile(base + _COMPILED_EXT):
            loader = importlib.machinery.SourcelessFileLoader(fullname, base + _COMPILED_EXT)
            return importlib.util.spec_from_file_location(fullname, base + _COMPILED_EXT, loader=loader)
        return None


def _resolve_root() -> str:
    """The live sources in the development container; on the GPU box the byte-compiled copy that
    ``oracle/fetch_ref.py`` built into ``oracle/_ref`` (code objects only; git-ignored, shipped by gpurun)."""
    env = os.environ.get("AH_REFERENCE_ROOT")
    if env:
        return env
    if _has_reference("/root/reference"):
        return "/root/reference"
    return _COMPILED_ROOT


REFERENCE_ROOT = _resolve_root()


def reference_available() -> bool:
    return _has_reference(REFERENCE_ROOT)


def reference_kind() -> str:
    """"source" (live .py tree) or "compiled" (sourceless modules built by oracle/fetch_ref.py)."""
    return "source" if os.path.isfile(os.path.join(REFERENCE_ROOT, "environment", "airhockey.py")) else "compiled"


_installed = False
_stubbed = {}


def uninstall() -> None:
    """Take the stub modules out of ``sys.modules`` again (the already-imported reference keeps its own references).
    A process that goes on to use torch must not be left with a ``MagicMock`` named ``tensorflow``: optional-dependency
    probes (``importlib.util.find_spec``) choke on a module without ``__spec__``."""
    global _installed
    for name, mod in _stubbed.items():
        if sys.modules.get(name) is mod:
            del sys.modules[name]
    _stubbed.clear()
    _installed = False


def install() -> None:
    """Install the stubs and put the reference on ``sys.path`` (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")

    class _Redis:  # no-op client standing in for redis.Redis / redis.StrictRedis
        def __init__(self, *a, **k):
            pass

        def get(self, key):
            return "null"

        def set(self, key, value):
            return True

        def publish(self, channel, payload):
            return 0

        def pubsub(self, **k):
            return MagicMock()

    redis = types.ModuleType("redis")
    redis.Redis = _Redis
    redis.StrictRedis = _Redis
    redis.ConnectionError = type("ConnectionError", (Exception,), {})
    sys.modules["redis"] = redis
    _stubbed["redis"] = redis

    for name in ("keras", "keras.backend", "keras.layers", "keras.models", "keras.optimizers",
                 "keras.initializers", "keras.utils", "tensorflow", "imp"):
        if name not in sys.modules:
            sys.modules[name] = _stubbed[name] = MagicMock()
    # `class NoisyDense(Layer)` must be definable
    sys.modules["keras.layers"].Layer = type("Layer", (object,), {})

    pytz = types.ModuleType("pytz")
    pytz.timezone = lambda name: datetime.timezone.utc
    if "pytz" not in sys.modules:
        sys.modules["pytz"] = _stubbed["pytz"] = pytz

    if not hasattr(np, "Infinity"):
        np.Infinity = np.inf

    if reference_kind() == "compiled":
        sys.meta_path.insert(0, _CompiledFinder(REFERENCE_ROOT))
    elif REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    os.environ.setdefault("PROJECT", tempfile.mkdtemp(prefix="ah_ref_project_"))
    logging.disable(logging.CRITICAL)
    _installed = True


def silence_side_channels() -> None:
    """No-op the Redis posts and the per-`done` CSV row (pure-simulation variant)."""
    install()
    import lib.rewards as rewards
    from lib.utils.connect import RedisConnection

    RedisConnection.post = lambda self, payload: None
    RedisConnection.publish = lambda self, channel, payload=True: None
    rewards.record_data_csv = lambda *a, **k: None


def numpy_dot_is_fma(trials: int = 20000, seed: int = 7) -> bool:
    """Self-probe: is ``np.dot([a,b],[c,d]) == fma(b, d, fl(a*c))`` on this build?

    Returns True when every discriminating case matches the FMA form, False when every one
    matches the plain form ``fl(fl(a*c) + fl(b*d))``; raises if the build is neither.
    """
    from fractions import Fraction

    rng = np.random.default_rng(seed)
    n_fma = n_plain = 0
    for _ in range(trials):
        a, b, c, d = (float(v) for v in rng.uniform(-35, 35, 4))
        plain = a * c + b * d
        fma = float(Fraction(b) * Fraction(d) + Fraction(a * c))
        if plain == fma:
            continue
        got = float(np.dot(np.array([a, b]), np.array([c, d])))
        n_fma += got == fma
        n_plain += got == plain
    if n_fma and not n_plain:
        return True
    if n_plain and not n_fma:
        return False
    raise RuntimeError(f"np.dot is neither plain nor fma-tail: fma={n_fma} plain={n_plain}")
