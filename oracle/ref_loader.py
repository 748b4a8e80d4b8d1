"""TEST INFRASTRUCTURE ONLY.  Import the UNMODIFIED reference module /root/reference/src/hashing.py.

Works only where /root/reference exists (the build container).  Missing third-party packages
(`datasketch`, `torch_geometric`) are satisfied by oracle/stubs; if the real ones are importable they
are used instead and `USING_STUBS` says so.  Nothing is copied out of /root/reference.
"""
import importlib
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get('SS_REFERENCE_ROOT', '/root/reference')
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stubs')
USING_STUBS = {}


def available():
    return os.path.exists(os.path.join(REFERENCE_ROOT, 'src', 'hashing.py'))


def load():
    """returns the reference's `src.hashing` module object"""
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    for name in ('datasketch', 'torch_geometric'):
        try:
            mod = importlib.import_module(name)
            USING_STUBS[name] = os.path.abspath(mod.__file__).startswith(_STUBS)
        except ImportError:
            USING_STUBS[name] = True
    if any(USING_STUBS.values()) and _STUBS not in sys.path:
        # real packages (if any) were imported above and stay in sys.modules; stubs fill the gaps
        sys.path.insert(0, _STUBS)
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', SyntaxWarning)
        return importlib.import_module('src.hashing')
