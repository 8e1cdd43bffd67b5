"""Build container only (@pytest.mark.ref: needs /root/reference): re-run the UNMODIFIED reference through
oracle/ref_loader.py exactly as tests/golden/make_golden.py does and require the committed fixtures to come out
bit-identical -- the golden vectors are the reference's outputs, not something edited by hand."""

import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_generator(tmp):
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(HERE, 'golden', 'make_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.HERE = str(tmp)            # write the regenerated fixtures next to nothing that is tracked
    return mod


def _same(a, b):
    if a.dtype.kind in 'fc':
        return np.array_equal(a, b, equal_nan=True)
    return np.array_equal(a, b)


@pytest.mark.ref
def test_fixtures_regenerate_bit_identically(tmp_path, capsys):
    from oracle import ref_loader
    mg = _load_generator(tmp_path)
    ref = ref_loader.load()
    simple = ('simple', {'slope': -1.0})
    with capsys.disabled():
        pass
    mg.sed_case(ref, 'fnu_mixed', 1500, 22, ('simple', {'slope': -0.7}), ('Calzetti_like', {}), 0.3, 6.8, observed=True)
    mg.sed_case(ref, 'lnu_dust', 2000, 21, simple, simple, 0.0, 7.0, observed=False)
    mg.lines_case(ref, 'dust', 1500, 41, simple, ('Calzetti_like', {}), 0.2, 7.0)
    mg.image_case(ref, 'adaptive8f', 2500, 34, 0.07, 64, ('adaptive', 8.))
    mg.image_case(ref, 'convgauss_odd', 2000, 36, 0.12, 41, ('convolved_gaussian', 0.35))
    mg.sed_index_case(ref)
    mg.observed_case(ref)
    made = sorted(f for f in os.listdir(tmp_path) if f.endswith('.npz'))
    assert len(made) >= 7
    for f in made:
        new = np.load(os.path.join(tmp_path, f), allow_pickle=False)
        old = np.load(os.path.join(HERE, 'golden', f), allow_pickle=False)
        assert sorted(new.files) == sorted(old.files), f
        for k in new.files:
            assert _same(new[k], old[k]), '%s: %s differs from the committed fixture' % (f, k)
