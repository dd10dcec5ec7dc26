"""A short, deterministic slice of tests/stress_gpu.py (random sizes / thresholds / seeds of the adversarial
generators through every operator, "nothing unexplained" comparators) so that the GPU suite itself exercises it;
run the script by hand for minutes of it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("round_", range(3))
def test_randomised_parity_slice(round_):
    import stress_gpu
    for k, case in enumerate(stress_gpu.CASES):
        seed = 900000 + 100 * round_ + k
        try:
            case(np.random.default_rng(seed))
        except BaseException as e:     # pytest.fail raises a BaseException subclass
            raise AssertionError("%s failed with seed %d: %r" % (case.__name__, seed, e)) from e
