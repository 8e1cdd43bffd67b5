"""Randomised small configurations (ragged sizes, empty inputs, duplicates, inf / NaN coordinates, extreme weights,
odd image sizes, every smoothing mode, 1..40 filters, missing dust arrays) through the public API against the oracle.
The generators live in tests/_fuzz_images.py and tests/_fuzz_sed.py (their run(seed, n_cases) take other seeds / more cases)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('seed', [0, 7])
def test_images_random_configurations_match_the_oracle(seed):
    from _fuzz_images import run
    assert run(seed, 40) == []


@pytest.mark.parametrize('seed', [0, 11])
def test_sed_random_configurations_match_the_oracle(seed):
    from _fuzz_sed import run
    assert run(seed, 30) == []
