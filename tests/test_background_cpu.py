"""CPU: host-side helpers of the static-background stage (fluidnexus_b200/background.py) against the restatements in
oracle/background_ref.py and closed forms.  The model / step classes themselves need a GPU (tests/test_background_gpu.py)."""
import math

import numpy as np
import pytest
import torch

from fluidnexus_b200 import background as B
from oracle import background_ref as OB


def test_quaternion_to_matrix_matches_build_rotation():
    q = torch.randn(64, 4, generator=torch.Generator().manual_seed(0)) * 3.0       # un-normalised on purpose
    R = B.quaternion_to_matrix(q)
    assert torch.allclose(R, OB.build_rotation(q), atol=1e-6)
    eye = torch.eye(3).expand(64, 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), eye, atol=1e-5)                   # rotations
    assert torch.allclose(torch.linalg.det(R), torch.ones(64), atol=1e-5)


def test_expon_lr_is_log_linear_with_eased_start():
    """FD/utils/general_utils.py:63-94: lr_init at step 0, lr_final at max_steps, geometric mean half way; with a delay the
    rate starts at lr_init * lr_delay_mult and eases in with a quarter sine."""
    f = B.expon_lr(1.6e-4, 1.6e-6, max_steps=30_000)
    assert f(0) == pytest.approx(1.6e-4) and f(30_000) == pytest.approx(1.6e-6) and f(10 ** 9) == pytest.approx(1.6e-6)
    assert f(15_000) == pytest.approx(math.sqrt(1.6e-4 * 1.6e-6))
    assert f(-1) == 0.0 and B.expon_lr(0.0, 0.0)(10) == 0.0
    g = B.expon_lr(1e-2, 1e-4, lr_delay_steps=100, lr_delay_mult=0.01, max_steps=1000)
    assert g(0) == pytest.approx(1e-2 * 0.01)
    t = 50 / 1000
    assert g(50) == pytest.approx((0.01 + 0.99 * math.sin(0.25 * math.pi)) * math.exp(math.log(1e-2) * (1 - t) + math.log(1e-4) * t))


def test_inverse_sigmoid_round_trip():
    x = torch.linspace(0.01, 0.99, 50)
    assert torch.allclose(torch.sigmoid(B.inverse_sigmoid(x)), x, atol=1e-6)
    assert torch.allclose(B.inverse_sigmoid(x), OB.inv_sigmoid(x))


def test_model_refuses_cpu():
    z = np.zeros((2, 3), np.float32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B.BackgroundModel(z, z, np.full((2, 1), 0.5, np.float32), z + 1, np.ones((2, 4), np.float32), device="cpu")
