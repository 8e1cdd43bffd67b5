"""CPU: the C-ABI library loads and exports every symbol include/synthobs_b200.h declares
(no compute calls without a GPU)."""

import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'synthobs_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(so_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from synthobs_b200 import _cabi
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(_cabi.lib, s), 'libsynthobs_b200.so does not export %s' % s
        assert s in _cabi.SIGNATURES, 'no ctypes signature for %s' % s
    assert sorted(_cabi.SIGNATURES) == syms
    assert b'sm_100a' in _cabi.lib.so_version()


def test_no_device_is_a_loud_error():
    import torch
    from synthobs_b200 import _cabi
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    assert _cabi.lib.so_device_ok() == 0
    with pytest.raises(_cabi.SynthObsCudaError):
        _cabi.require_device()
    import numpy as np
    from synthobs_b200.morph import images
    with pytest.raises(_cabi.SynthObsCudaError):
        images.core(np.zeros(3), np.zeros(3), np.ones(3))


def test_product_does_not_import_oracle():
    """the product path must never route through oracle/ (or any CPU fallback)"""
    pkg = os.path.join(ROOT, 'synthobs_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py') or f.endswith('.cu') or f.endswith('.cuh'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f
