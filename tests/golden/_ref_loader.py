"""Import the UNMODIFIED reference hot-path modules from /root/reference (container only).

Test infrastructure.  Used by ``make_golden.py`` (fixture generation) and by the
``needs_reference`` tests that cross-check the oracle against the live reference
when /root/reference is mounted.  Nothing on the product path, in ``-m gpu``
tests, ``smoke()`` or ``bench.py`` imports this file: /root/reference does not
exist on the GPU box.

``import a2c`` fails in this image (a2c/__init__.py:1 eagerly imports logger ->
matplotlib, preprocessing -> skimage, runner -> gym / mlagents_envs / gym_unity /
ml_utils; none installed).  We register a bare ``a2c`` namespace whose
``__path__`` points at the reference and stub exactly the third-party names the
files import at top level.  No reference file is edited or
copied.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("A2C_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "a2c", "updater.py"))


def _stub(name, **attrs):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__path__ = []  # behave like a package so dotted imports resolve
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def load_reference():
    """Returns a namespace with .utils .preprocessing .models .updater .runner."""
    if not reference_available():
        raise RuntimeError("reference tree not mounted at %s" % REF_ROOT)
    if "a2c" not in sys.modules or not hasattr(sys.modules["a2c"], "_graft_stub"):
        pkg = types.ModuleType("a2c")
        pkg.__path__ = [os.path.join(REF_ROOT, "a2c")]
        pkg._graft_stub = True
        sys.modules["a2c"] = pkg

        def _try_key(d, k, default):  # ml_utils.utils.try_key (same as a2c/utils.py:4-7)
            return d[k] if k in d else default

        _stub("skimage")
        _stub("skimage.color", rgb2grey=lambda x: x)
        _stub("gym")
        _stub("ml_utils")
        _stub("ml_utils.utils", try_key=_try_key)
        _stub("mlagents_envs")
        _stub("mlagents_envs.environment", UnityEnvironment=object)
        _stub("mlagents_envs.side_channel")
        _stub("mlagents_envs.side_channel.engine_configuration_channel",
              EngineConfigurationChannel=object)
        _stub("mlagents_envs.side_channel.environment_parameters_channel",
              EnvironmentParametersChannel=object)
        _stub("gym_unity")
        _stub("gym_unity.envs", UnityToGymWrapper=object)
    ns = types.SimpleNamespace()
    for sub in ("utils", "preprocessing", "models", "updater", "runner"):
        setattr(ns, sub, importlib.import_module("a2c." + sub))
    return ns
