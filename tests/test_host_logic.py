"""Host-side logic that needs no GPU: synthetic configs[3] inputs, gradient-bucket plumbing switches, ABI struct
layouts the Python side builds by hand."""
import ctypes

import torch

from eda_b200 import _lib, attn_ops, ddp, hotpath


def test_synthetic_inputs_have_configs3_shapes():
    pc, pos, text, text_mask, det, det_mask, query, qpos = hotpath.synthetic_inputs(2, N=4096)
    assert pc.shape == (2, 4096, 6) and pos.shape == (2, 1024, hotpath.D_MODEL)
    assert text.shape == (2, 80, 288) and det.shape == (2, 132, 288) and query.shape == (2, 256, 288)
    assert qpos.shape == (2, 256, 6) and (qpos[..., 3:] > 0).all()          # box sizes are positive
    assert text_mask.dtype == torch.bool and text_mask.shape == (2, 80) and det_mask.shape == (2, 132)
    assert not text_mask[0].any() and not det_mask[0].any()                  # row 0 keeps every token (L = 80 exactly)
    keep = (~text_mask).sum(1)
    assert (keep >= 20).all() and ((~text_mask).long().cumsum(1) == torch.arange(1, 81).clamp(max=keep[:, None])).all()
    again = hotpath.synthetic_inputs(2, N=4096)
    assert all(torch.equal(a, b) for a, b in zip((pc, pos, text, query), (again[0], again[1], again[2], again[6])))


def test_flat_gradients_on_cpu_tags_nothing():
    m = torch.nn.Linear(4, 3)
    fg = ddp.FlatGradients(m)
    assert not fg.fused                              # only a CUDA bucket turns fused weight gradients on
    assert not any(attn_ops.fused_grad_enabled(p) for p in m.parameters())
    assert fg.check_views() and fg.flat.numel() == 15
    m(torch.ones(2, 4)).sum().backward()
    fg.sync()                                        # no side stream on CPU: a no-op
    assert torch.equal(m.bias.grad, torch.full((3,), 2.0)) and fg.flat.abs().sum() > 0
    attn_ops.join_wgrad()                            # nothing pending: must not touch CUDA
    fg.zero()
    assert fg.flat.abs().sum() == 0 and fg.check_views()
    fg.release()                                     # CPU bucket: nothing to switch off, must not raise


def test_fused_grad_buffers_are_scoped_to_the_owning_bucket():
    """fused weight-gradient accumulation must not be a process-wide switch.  Only parameters tagged by a
    live FlatGradients whose `fused` flag is on hand their .grad to the kernels; other models in the process, released
    buckets and dropped buckets go through autograd."""
    import gc

    class Owner:  # stands in for a CUDA FlatGradients (no GPU in this suite): the tag protocol is what is tested
        fused = True

    import weakref
    p, q = torch.nn.Parameter(torch.zeros(3, 3)), torch.nn.Parameter(torch.zeros(3, 3))
    p.grad, q.grad = torch.zeros(3, 3), torch.zeros(3, 3)
    assert attn_ops._grad_buffers((p, q, None)) == [None, None, None]        # nobody owns them
    owner = Owner()
    p._eda_fused_grad_owner = weakref.ref(owner)
    got = attn_ops._grad_buffers((p, q, None, torch.nn.Parameter(torch.zeros(2))))
    assert got[0] is p.grad and got[1] is None and got[2] is None and got[3] is None  # q belongs to "another model"
    owner.fused = False                                                       # release()
    assert attn_ops._grad_buffers((p,)) == [None]
    owner.fused = True
    del owner
    gc.collect()
    assert attn_ops._grad_buffers((p,)) == [None]                             # bucket dropped without release()


def test_pack_registry_is_per_model():
    """the packed-weight registry of a GraphedTrainStep is attached to ITS model's modules."""
    a, b = torch.nn.Linear(4, 4), torch.nn.Linear(4, 4)
    reg = attn_ops.PackRegistry(torch.device("cpu"))
    reg.attach(a)
    assert attn_ops.active_registry(a) is None       # attached but not recording
    reg.active = True
    assert attn_ops.active_registry(a) is reg and attn_ops.active_registry(b) is None
    reg.active = False
    reg.detach(a)
    assert "_eda_pack_registry" not in a.__dict__


def test_hand_built_abi_structs_match_the_header_layout():
    # attn_ops.PackRegistry writes descriptor rows as 6 int64 (w, dst, stride_n, stride_k, N | K << 32, Kpad):
    # struct eda_linear_pack_desc { const float *w; float *dst; long long stride_n, stride_k; int N, K, Kpad, reserved; }
    class Desc(ctypes.Structure):
        _fields_ = [("w", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("sn", ctypes.c_longlong), ("sk", ctypes.c_longlong),
                    ("N", ctypes.c_int), ("K", ctypes.c_int), ("Kpad", ctypes.c_int), ("reserved", ctypes.c_int)]

    assert ctypes.sizeof(Desc) == 48
    row = (ctypes.c_longlong * 6)(0x1000, 0x2000, 288, 1, 288 | (256 << 32), 256)
    d = Desc.from_buffer_copy(bytes(row))
    assert (d.w, d.dst, d.sn, d.sk, d.N, d.K, d.Kpad, d.reserved) == (0x1000, 0x2000, 288, 1, 288, 256, 256, 0)
    assert ctypes.sizeof(_lib.LinearProblem) == 6 * 8 + 4 * 4 + 8 + 2 * 4  # 6 pointers, 4 ints, pre_ln, y_row_stride + reserved
    assert ctypes.sizeof(_lib.WgradProblem) == 4 * 8 + 8 + 3 * 4 + 4 + 2 * 8  # 4 pointers, rows, 3 ints (+pad), 2 pointers
