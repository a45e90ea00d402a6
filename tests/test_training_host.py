"""CPU: host logic of the training-step forward (panfusion_b200/training.py) and its oracle (oracle/training.py;
models/pano/PanFusion.py:78-97)."""
import pytest
import torch


def test_schedule_and_add_noise_restatement():
    """The product's schedule table is the oracle's (SD-2 scheduler config: scaled_linear 0.00085..0.012, 1000 steps), and the
    oracle's add_noise is the published rule: per-sample sqrt(abar_t) x0 + sqrt(1 - abar_t) eps; t = 0 is nearly clean, t = 999
    nearly pure noise."""
    from oracle import sampler as osamp, training as otr
    from panfusion_b200.sampler import DDIMSchedule
    abar = osamp.DDIM().alphas_cumprod
    assert torch.equal(DDIMSchedule().alphas_cumprod, abar)
    g = torch.Generator().manual_seed(0)
    x0, eps = torch.randn(3, 2, 4, 8, 8, generator=g), torch.randn(3, 2, 4, 8, 8, generator=g)
    t = torch.tensor([0, 500, 999])
    z = otr.add_noise(x0, eps, t, abar)
    for i in range(3):
        a = abar[t[i]].double()
        want = a.sqrt() * x0[i].double() + (1 - a).sqrt() * eps[i].double()
        torch.testing.assert_close(z[i].double(), want, rtol=1e-6, atol=1e-6)
    assert (z[0] - x0[0]).abs().max() < 0.15 and (z[2] - eps[2]).abs().max() < 0.35


def test_training_step_needs_cuda_and_has_no_backward():
    from panfusion_b200.training import TrainingStep
    step = TrainingStep(mv_base_model=None)
    with pytest.raises(NotImplementedError):
        step.training_step()
    x = torch.zeros(1, 4)
    with pytest.raises(Exception):  # no CPU path: host tensors are refused by the CUDA wrapper
        step.add_noise(x, x, torch.zeros(1, dtype=torch.int64))
