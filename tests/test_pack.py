"""Packed input format: the native batch packer against outputs of the reference's own
`SVGTensorDataset.get_data` (tests/golden/make_pack_golden.py), and the GPU unpack kernel."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _fixture():
    fx = dict(np.load(os.path.join(HERE, "golden", "pack_batch.npz"), allow_pickle=False))
    G = int(fx["G"])
    rows, off = torch.from_numpy(fx["rows"]), fx["offsets"]
    n = (len(off) - 1) // G
    icons = []
    for i in range(n):
        paths = []
        for g in range(G):
            r0, r1 = int(off[i * G + g]), int(off[i * G + g + 1])
            if r1 > r0:
                paths.append(rows[r0:r1].clone())
        icons.append(paths)
    return fx, icons


def test_native_packer_matches_reference_get_data():
    from deepsvg_b200 import pack_icons
    fx, icons = _fixture()
    G, S, TOTAL = int(fx["G"]), int(fx["S"]), int(fx["TOTAL"])
    pb = pack_icons(icons, G, S, grouped=False, pin=False)
    assert pb.cmd.dtype == torch.uint8 and pb.args.dtype == torch.int16
    assert np.array_equal(pb.cmd.numpy().astype(np.float32), fx["want_commands"])
    assert np.array_equal(pb.args.numpy().astype(np.float32), fx["want_args"])
    pg = pack_icons(icons, G, TOTAL, grouped=True, pin=False)
    assert pg.cmd.shape == (len(icons), 1, TOTAL + 2)
    assert np.array_equal(pg.cmd.numpy().astype(np.float32), fx["want_commands_grouped"])
    assert np.array_equal(pg.args.numpy().astype(np.float32), fx["want_args_grouped"])
    # 23 bytes per position instead of 48
    assert pb.nbytes == fx["want_commands"].size + 2 * fx["want_args"].size


def test_packer_rejects_what_the_reference_cannot_stack():
    from deepsvg_b200 import pack_icons, pack_tensors
    long_path = torch.zeros(31, 14)
    long_path[:, 1:] = -1
    with pytest.raises(RuntimeError, match="holds 30"):
        pack_icons([[long_path]], 8, 30, pin=False)
    bad = torch.zeros(2, 14)
    bad[1, 12] = 3.5
    with pytest.raises(RuntimeError, match="not an integer"):
        pack_icons([[bad]], 8, 30, pin=False)
    with pytest.raises(ValueError):
        pack_icons([[torch.zeros(1, 14)] * 9], 8, 30, pin=False)
    c = torch.tensor([[[5.0, 0.0, 4.0]]])
    a = torch.full((1, 1, 3, 11), -1.0)
    pb = pack_tensors(c, a, pin=False)
    assert pb.cmd.tolist() == [[[5, 0, 4]]] and int(pb.args.min()) == -1
    with pytest.raises(ValueError):
        pack_tensors(c + 0.5, a, pin=False)


def test_unpack_has_no_cpu_path():
    from deepsvg_b200 import pack_icons
    _, icons = _fixture()
    with pytest.raises(RuntimeError, match="no CPU path"):
        pack_icons(icons, 8, 30, pin=False).unpack()


@pytest.mark.gpu
def test_gpu_unpack_and_forward_from_a_packed_batch():
    """packed host batch -> H2D -> unpack kernel == the float tensors; the model output from them is identical."""
    from deepsvg_b200 import Hierarchical, SVGTransformer, pack_icons
    fx, icons = _fixture()
    pb = pack_icons(icons, 8, 30, labels=list(range(len(icons))))
    dev = pb.cuda()
    c, a = dev.unpack()
    assert np.array_equal(c.cpu().numpy(), fx["want_commands"]) and np.array_equal(a.cpu().numpy(), fx["want_args"])
    assert torch.equal(dev.label.cpu(), torch.arange(len(icons)))
    model = SVGTransformer(Hierarchical(use_vae=False)).cuda().eval()
    with torch.no_grad():
        o1 = model(c, a, c, a)["args_logits"]
        cf, af = torch.from_numpy(fx["want_commands"]).cuda(), torch.from_numpy(fx["want_args"]).cuda()
        o2 = model(cf, af, cf, af)["args_logits"]
    assert torch.equal(o1, o2)
