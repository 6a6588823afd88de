"""Device-side token plumbing (sa_tokens_prepare / gather / narrow) against the host-side prepare_batch mirror."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ordering(grid):
    from synthanatomy_b200.networks.transformers import Ordering
    return Ordering("raster_scan", 3, (1, *grid), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose"))


@pytest.mark.parametrize("grid,dtype", [((10, 14, 10), torch.uint16), ((10, 14, 10), torch.int64), ((20, 28, 25), torch.uint16),
                                        ((3, 5, 2), torch.int32)])
def test_prepare_batch_device_equals_host_prepare_batch(grid, dtype):
    from synthanatomy_b200.utils import tokens as tk
    from synthanatomy_b200.utils.transformer import prepare_batch
    order = _ordering(grid)
    g = torch.Generator().manual_seed(5)
    q = torch.randint(0, 2048, (6, *grid), generator=g)
    (x_ref, _), y_ref = prepare_batch({"quantization": q}, order.get_sequence_ordering(), 2048)
    dev = tk.DeviceOrdering(order, "cuda")
    stored = torch.from_numpy(q.numpy().astype({torch.uint16: np.uint16, torch.int32: np.int32, torch.int64: np.int64}[dtype]))
    (x, cond), y = tk.prepare_batch_device({"quantization": stored}, dev, 2048)
    assert cond is None and x.dtype == torch.int64 and y.dtype == torch.int64
    assert torch.equal(x.cpu(), x_ref) and torch.equal(y.cpu(), y_ref)
    # the sampling tail: sequence -> grid is the inverse gather
    back = tk.sequence_to_grid(y, dev, order.dimensions)
    assert tuple(back.shape) == (6, *grid) and torch.equal(back.cpu(), q)


def test_prepare_batch_device_conditionings_and_empty_batch():
    from synthanatomy_b200.utils import tokens as tk
    order = _ordering((3, 5, 2))
    dev = tk.DeviceOrdering(order, "cuda")
    q = torch.randint(0, 2048, (4, 3, 5, 2))
    (x, cond), y = tk.prepare_batch_device({"quantization": q, "age": torch.tensor([3, 1, 4, 1])}, dev, 2048, ("age",))
    assert tuple(cond[0].shape) == (4, 1) and cond[0].is_cuda and cond[0].dtype == torch.int64
    assert int(x[:, 0].min()) == 2048 and int(x[:, 0].max()) == 2048
    (x0, _), y0 = tk.prepare_batch_device({"quantization": q[:0]}, dev, 2048)
    assert tuple(x0.shape) == (0, 30) and tuple(y0.shape) == (0, 30)


def test_tokens_narrow_and_save_from_device(tmp_path):
    from synthanatomy_b200 import pf_ops as pf
    from synthanatomy_b200.utils import tokens as tk
    idx = torch.randint(0, 2048, (2, 10, 14, 10), device="cuda")
    u16 = pf.tokens_narrow(idx)
    assert u16.dtype == torch.uint16 and torch.equal(u16.cpu().to(torch.int64), idx.cpu())
    with pytest.raises(ValueError):
        pf.tokens_narrow(torch.tensor([5, 65536], device="cuda"))
    with pytest.raises(ValueError):
        pf.tokens_narrow(torch.tensor([-1], device="cuda"))
    paths = tk.save_token_volumes(idx, ["a.nii.gz", "b.nii.gz"], str(tmp_path))
    assert np.array_equal(tk.load_token_volume(paths[1]), idx[1].cpu().numpy().astype(np.uint16))
