"""Drop-in boundary: every class / function of the reference's public API on the hot path exists in
this package under the same import path with the same parameter names, in the same order, with defaults where the
reference has them.  The reference's side is a snapshot taken from the LIVE reference
(tests/golden/api_signatures.json, oracle/make_api_snapshot.py).  CPU only."""
import importlib
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SNAP = json.load(open(os.path.join(HERE, 'golden', 'api_signatures.json')))


def _resolve(key):
    parts = key.split('.')
    for cut in range(len(parts), 0, -1):
        try:
            obj = importlib.import_module('.'.join(parts[:cut]))
        except ImportError:
            continue
        for name in parts[cut:]:
            obj = getattr(obj, name)
        return obj
    raise ImportError(key)


def _params(fn):
    sig = inspect.signature(fn)
    return [[n, p.default is not inspect.Parameter.empty] for n, p in sig.parameters.items()
            if p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)]


@pytest.mark.parametrize('key', sorted(k for k, v in SNAP.items() if isinstance(v, list)))
def test_same_parameters_as_the_reference(key):
    ours = _params(_resolve(key))
    ref = SNAP[key]
    # same leading parameters (names and order); this package may append optional ones
    assert [n for n, _ in ours[:len(ref)]] == [n for n, _ in ref], (ours, ref)
    for (n, has_default), (_, ref_default) in zip(ours, ref):
        assert has_default or not ref_default, 'parameter %s lost its default' % n
    for n, has_default in ours[len(ref):]:
        assert has_default, 'extra parameter %s must be optional' % n


def test_the_module_is_this_package_not_the_reference():
    import diff_gpmp2
    import dgpmp2_b200
    assert os.path.dirname(os.path.abspath(diff_gpmp2.__file__)).startswith(os.path.dirname(HERE))
    from diff_gpmp2.gpmp2 import PlanLayer
    assert PlanLayer.__module__.startswith('dgpmp2_b200') or 'dgpmp2_b200' in inspect.getsourcefile(PlanLayer)
    assert dgpmp2_b200 is not None
