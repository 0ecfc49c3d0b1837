"""Host logic: the Hydra-compatible composer reproduces the reference's config tree and override syntax."""
import pytest

from opticalflowdiffusion_b200.config import compose


def test_defaults_match_reference_values():
    c = compose()
    a = c.algorithm
    assert a.name == "flow_diffuser" and a.target == "joint" and a.timesteps == 1000 and a.flow_max == 20
    assert a.lr == pytest.approx(1e-5) and a.weight_decay == pytest.approx(1e-6) and isinstance(a.lr, float)
    assert a.zero_init is True and a.latent is False and a.noiser == "image" and a.image_size == 128
    e = c.experiment
    assert e.name == "matrix_flow" and e.training.data.batch_size == 16 and e.training.clipping == 100
    assert e.training.precision == 32 and e.training.optim.accumulate_grad_batches == 1      # inherited from base
    assert e.validation.data.batch_size == 8 and e.validation.limit_batch == 1


def test_overrides():
    c = compose(["algorithm.target=flow", "algorithm.sampling_timesteps=50", "+algorithm.cond_channels=6",
                 "algorithm.image_size=[436,1024]", "experiment.tasks=[validation]"])
    assert c.algorithm.target == "flow" and c.algorithm.sampling_timesteps == 50 and c.algorithm.cond_channels == 6
    assert c.algorithm.image_size == [436, 1024] and c.experiment.tasks == ["validation"]
    with pytest.raises(KeyError):
        compose(["algorithm.does_not_exist=1"])
    with pytest.raises(FileNotFoundError):
        compose(["algorithm=pwc_learner"])      # outside the hot path


def test_state_dict_names_match_reference_checkpoint_layout():
    """841 keys: the 276 UNet tensors under unet.*, _model.*, model.model.* + 13 schedule buffers."""
    import torch
    from opticalflowdiffusion_b200 import FlowDiffuser
    for target, n_model_prefix in (("flow", "model.model."), ("joint", "model.model.model.")):
        m = FlowDiffuser(compose([f"algorithm.target={target}"]).algorithm)
        sd = m.state_dict()
        assert len(sd) == 841
        assert "unet.downs.0.2.fn.fn.to_out.1.g" in sd and n_model_prefix + "final_conv.bias" in sd
        assert sum(k.startswith("model.") and "." not in k[6:] for k in sd) == 13
    # zero_init zeroes final_conv for the warp targets only (flow_diffuser.py:31-33)
    assert float(sd["unet.final_conv.weight"].abs().sum()) == 0.0


def test_schedule_and_ddim_grid_match_golden(golden):
    import numpy as np
    from opticalflowdiffusion_b200.diffusion import ConditionalDiffusion, ddim_time_pairs
    import torch
    g = golden("schedule")
    d = ConditionalDiffusion(torch.nn.Identity(), 64, timesteps=1000)
    for k in ConditionalDiffusion.BUFFERS:
        assert np.array_equal(getattr(d, k).numpy(), g[k]), k
    for T, S in ((1000, 50), (1000, 7), (50, 10), (6, 3)):
        times = g[f"ddim_times_{T}_{S}"].tolist()
        assert ddim_time_pairs(T, S) == list(zip(times[:-1], times[1:]))
