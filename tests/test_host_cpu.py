"""Host-side logic and the C-ABI surface; no GPU needed."""
import ctypes
import pathlib
import re

import numpy as np
import pytest
import torch

import spcl_b200
from spcl_b200 import _native as nat
from spcl_b200 import ops
from spcl_b200.losses import _pick_tc
from oracle.closed_form import codes_from_target

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _same_classes(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.array_equal(a[:, None] == a[None, :], b[:, None] == b[None, :])


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "spcl.h").read_text()
    declared = set(re.findall(r"\b(spcl_[a-z0-9_]+)\s*\(", header))
    declared.discard("spcl_error_string() ")
    assert declared == set(nat.ALL_SYMBOLS), declared ^ set(nat.ALL_SYMBOLS)
    handle = nat.lib()
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.spcl_version() >= 100
    assert handle.spcl_error_string(0) == b"ok"


def test_abi_validates_arguments_without_touching_the_gpu():
    h = nat.lib()
    null = ctypes.c_void_p(0)
    rc = h.spcl_supcon_fwd_f32(null, 128, 64, 64, null, null, 64, 0, 128, 14.0, 1.0, 0, null, 128, null, null)
    assert rc == -1
    rc = h.spcl_supcon_finalize(null, 0, 0, null, null)
    assert rc == -1
    fake = ctypes.c_void_p(4096)
    # d > SPCL_MAX_D is rejected before any launch
    rc = h.spcl_supcon_fwd_f32(fake, 128, 512, 512, fake, null, 64, 0, 128, 14.0, 1.0, 0, fake, 128, fake, null)
    assert rc == -2
    # row_begin must be tile aligned on the tensor-core path
    rc = h.spcl_supcon_fwd_bf16(fake, 256, 256, 128, fake, fake, 64, 256, 14.0, 1.0, 0, fake, fake, fake, null)
    assert rc == -2
    assert b"invalid" in h.spcl_error_string(-1)


@pytest.mark.parametrize("target", [
    [0, 1, 2, 0, 1, 2], [5, 5, 5, 5, 5, 5], [2 ** 24, 2 ** 24 + 1, 7, 7, 2 ** 24 + 2, 0],
    [0.5, 0.25, 0.5, 1.0, 1.0, 0.25], [-3, 3, -3, 3, 0, 0],
])
def test_label_codes_lists_match_reference_semantics(target):
    codes = ops.label_codes(target, len(target), "cpu").numpy()
    assert codes.dtype == np.int32
    assert _same_classes(codes, codes_from_target(target))
    # and against the reference's literal construction (contrast_loss3.py:135-136)
    t = torch.Tensor(target)
    assert np.array_equal((t[:, None] == t[None, :]).numpy(), codes[:, None] == codes[None, :])


@pytest.mark.parametrize("dtype", [torch.int64, torch.int32, torch.uint8, torch.float32])
def test_label_codes_tensors(dtype):
    t = torch.tensor([3, 1, 3, 0, 1, 200], dtype=dtype)
    codes = ops.label_codes(t, 6, "cpu").numpy()
    assert _same_classes(codes, t.numpy())
    big = torch.tensor([2 ** 40, 2 ** 40 + 1, 2 ** 40], dtype=torch.int64)
    assert _same_classes(ops.label_codes(big, 3, "cpu").numpy(), big.numpy())


def test_label_codes_shape_errors():
    with pytest.raises(AssertionError):
        ops.label_codes([0, 1, 2], 4, "cpu")
    with pytest.raises(TypeError):
        ops.label_codes("abc", 3, "cpu")


def test_tri_codes():
    m = torch.tensor([[1.0, 0.0, 0.5], [2.0, 1.0, 0.0], [0.0, -1.0, 1.0]])
    out = ops.tri_codes(m, 3, "cpu")
    assert out.tolist() == [[1, 0, 2], [2, 1, 0], [0, 2, 1]]


def test_precision_policy():
    assert _pick_tc("bf16", 128, False) and not _pick_tc("fp32", 10 ** 6, False)
    assert not _pick_tc("auto", 512, False) and _pick_tc("auto", 1024, False)
    assert not _pick_tc("auto", 10 ** 5, True)        # tri-state mask -> fp32 path
    with pytest.raises(ValueError):
        _pick_tc("fp8", 128, False)


def test_module_surface_matches_reference():
    sp = spcl_b200.SelfPacedSupConLoss(temperature=0.1, weight_update="soft", correct_grad=True)
    assert sp.age_param == 1e6                       # contrast_loss3.py:122
    sp.set_gamma(3)
    assert sp.age_param == 3.0 and isinstance(sp.age_param, float)
    assert repr(sp) == "SelfPacedSupConLoss with T: 0.1, method: soft gamma: 3.0"
    s1 = spcl_b200.SupConLoss1()
    assert s1._t == 0.07
    sx = spcl_b200.SupConLoss1(exclude_other_pos=True)          # :35, :97-100
    assert sx._exclude_pos is True and sx._gamma_mode_cg()[1] == spcl_b200._native.MODE_EXCL
    with pytest.raises(spcl_b200._native.SpclError):            # exclude_other_pos is an fp32-path mode
        spcl_b200.losses._pick_tc("bf16", 4096, False, spcl_b200._native.MODE_EXCL)
    assert spcl_b200.losses._pick_tc("auto", 1 << 15, False, spcl_b200._native.MODE_EXCL) is False
    with pytest.raises(AttributeError):
        _ = sp.downgrade_ratio                       # only after a forward call


def test_no_cpu_fallback():
    z = torch.nn.functional.normalize(torch.randn(8, 16), dim=1)
    crit = spcl_b200.SupConLoss1()
    with pytest.raises(RuntimeError, match="no CPU path"):
        crit(z, z.clone(), target=[0, 1] * 4)
    with pytest.raises(AssertionError):             # the reference's shape assert (:155)
        crit(z, z[:4], target=[0, 1] * 4)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(nat.SpclError, match="no CPU/PyTorch fallback"):
        nat.lib()


def test_pscheduler_follows_reference_formula():
    s = spcl_b200.PScheduler(max_epoch=80, begin_value=2.0, end_value=60.0, p=0.5)
    vals = []
    for _ in range(5):
        vals.append(s.value)
        s.step()
    expect = [2.0 + 58.0 * (e / 80) ** 0.5 for e in range(5)]     # semi_seg/hooks/infonce.py:50-53
    assert np.allclose(vals, expect)
    st = s.state_dict()
    s2 = spcl_b200.PScheduler(max_epoch=80, begin_value=2.0, end_value=60.0, p=0.5)
    s2.load_state_dict(st)
    assert s2.value == s.value


def test_workload_shapes():
    from spcl_b200.workloads import WORKLOADS, make_workload
    z1, z2, lab = make_workload("cfg1_cpu_2x64_d128")
    assert z1.shape == (64, 128) and lab.shape == (64,)
    assert torch.allclose(z1.norm(dim=1), torch.ones(64), atol=1e-5)
    assert set(WORKLOADS) >= {"cfg2_encoder_2x256_d256", "cfg3_dense_2x16384_d128_simclr"}


# ---- symmetric pass A: the tile enumeration the kernel walks (host copy of the same code) --------------------
@pytest.mark.parametrize("RB,bn", [(1, 256), (2, 256), (3, 128), (5, 256), (8, 128), (16, 256), (37, 256), (64, 128),
                                   (256, 256), (255, 256)])
@pytest.mark.parametrize("vg", [1, 2, 7, 148, 296])
def test_symmetric_tile_walk_covers_the_triangle_exactly_once(RB, bn, vg):
    """Every (row block I, column tile t >= first tile of I) is visited by exactly one CTA of the (virtual) grid --
    also when the grid is the concatenation of several ranks' launches -- ranges are balanced to one tile, and the
    segment flags (`first` / `last`: A-tile reload, row flush) bracket each row-block segment of a CTA."""
    h = nat.lib()
    h.spcl_debug_sym_walk.restype = ctypes.c_int64
    h.spcl_debug_sym_walk.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_void_p, ctypes.c_int64]
    per128 = bn // 128
    CT = -(-RB // per128)
    want = {(I, t) for I in range(RB) for t in range(I // per128, CT)}
    seen = {}
    counts = []
    cap = len(want) + 8
    buf = np.zeros((cap, 4), dtype=np.int32)
    for vb in range(vg):
        k = h.spcl_debug_sym_walk(RB, CT, bn, vb, vg, buf.ctypes.data_as(ctypes.c_void_p), cap)
        assert 0 <= k <= cap
        counts.append(k)
        tiles = buf[:k].copy()
        for idx, (I, t, first, last) in enumerate(tiles):
            assert (I, t) not in seen, f"tile {(I, t)} visited by CTA {seen.get((I, t))} and {vb}"
            seen[(int(I), int(t))] = vb
            starts = idx == 0 or tiles[idx - 1][0] != I
            ends = idx == k - 1 or tiles[idx + 1][0] != I
            assert bool(first) == starts and bool(last) == ends, (vb, idx, I, t, first, last)
            if not starts:
                assert t == tiles[idx - 1][1] + 1          # consecutive column tiles inside a segment
    assert set(seen) == want
    assert max(counts) - min(counts) <= 1
    assert h.spcl_debug_sym_walk(RB, CT, bn, vg, vg, None, 0) == -1        # vb out of range


# ---- hook-side glue and dense front end (8 f4): host logic --------------------------------------
def test_encode_labels_equals_sklearn_label_encoder():
    from spcl_b200.hooks import encode_labels
    cases = [["b", "a", "c", "a"], ["patient010", "patient002", "patient010", "patient100"], [3, 1, 2, 1, 3],
             ["0", "1", "2"] * 4]
    assert encode_labels(cases[0]) == [1, 0, 2, 0]
    assert encode_labels(cases[1]) == [1, 0, 1, 2]
    sk = pytest.importorskip("sklearn.preprocessing")
    for v in cases:
        assert encode_labels(v) == sk.LabelEncoder().fit(v).transform(v).tolist()      # helper.py:48-55


def test_get_label_dispatch_follows_reference():
    from spcl_b200.hooks import get_label
    part = ["0", "1", "2", "0", "1", "2"]
    group = ["patient003_00", "patient003_00", "patient003_01", "patient001_00", "patient001_01", "patient001_01"]
    assert get_label("partition", "acdc", part, group) == [0, 1, 2, 0, 1, 2]
    assert get_label("patient", "acdc", part, group) == [1, 1, 1, 0, 0, 0]
    assert get_label("cycle", "acdc", part, group) == [0, 0, 1, 0, 1, 1]
    assert get_label("self", "acdc", part, group) == list(range(6))
    assert get_label("patient", "prostate", part, group) == [1, 1, 1, 0, 0, 0]
    with pytest.raises(NotImplementedError):
        get_label("cycle", "prostate", part, group)           # hooks/utils.py:23-31: no cycle labels there
    with pytest.raises(NotImplementedError):
        get_label("partition", "unknown", part, group)


def test_device_meter_and_label_cache():
    from spcl_b200.hooks import DeviceMeter, LabelCache
    m = DeviceMeter()
    assert m.summary() != m.summary()                          # nan when empty
    for v in (torch.tensor(1.0), torch.tensor(2.0), 6.0):
        m.add(v)
    assert m.summary() == pytest.approx(3.0)
    m.reset()
    m.add(4.0)
    assert m.summary() == 4.0
    cache = LabelCache(capacity=2)
    a = cache([0, 1, 2], "cpu")
    assert a.dtype == torch.int32 and a.tolist() == [0, 1, 2]
    assert cache([0, 1, 2], "cpu") is a                        # second call: no new upload
    cache([1], "cpu"); cache([2], "cpu")
    assert cache([0, 1, 2], "cpu") is not a                    # evicted (LRU, capacity 2)


def test_self_paced_gamma_schedule_steps_like_the_hook_factory():
    from spcl_b200.hooks import SelfPacedGammaSchedule
    s = SelfPacedGammaSchedule(mode="soft", p=0.5, begin_value=5.0, end_value=60.0, max_epoch=4, correct_grad=True)
    seen = [s.new_epoch() for _ in range(4)]
    assert seen == pytest.approx([5.0 + 55.0 * (e / 4) ** 0.5 for e in range(4)])       # infonce.py:46-49, :134-136
    assert s.criterion.age_param == pytest.approx(seen[-1])


def test_dense_abi_validates_arguments_without_touching_the_gpu():
    h = nat.lib()
    null, fake = ctypes.c_void_p(0), ctypes.c_void_p(4096)
    assert h.spcl_dense_rows_fwd(null, null, 2, 8, 16, 16, 4, 4, 16, 1e-12, fake, fake, null) == -1
    assert h.spcl_dense_rows_fwd(fake, null, 2, 8, 16, 16, 0, 4, 16, 1e-12, fake, fake, null) == -1
    assert h.spcl_dense_rows_fwd(fake, null, 2, 8, 16, 16, 32, 4, 128, 1e-12, fake, fake, null) == -2   # pools up
    assert h.spcl_dense_rows_fwd(fake, fake, 2, 8, 16, 16, 4, 4, 0, 1e-12, fake, fake, null) == -1      # P = 0
    assert h.spcl_dense_rows_bwd(null, null, 2, 8, 16, 16, 4, 4, 16, fake, null) == -1
    assert h.spcl_dense_rows_bwd(fake, null, 2, 8, 16, 16, 4, 32, 128, fake, null) == -2


def test_dense_rows_rejects_cpu_tensors_and_bad_points():
    import spcl_b200
    with pytest.raises(RuntimeError):
        spcl_b200.ops.dense_rows(torch.randn(1, 4, 8, 8), (4, 4))                   # no CPU path
    with pytest.raises(ValueError):
        spcl_b200.ops.dense_rows(torch.randn(4, 8, 8), (4, 4))


def test_problem_struct_matches_the_header(tmp_path):
    """ctypes mirror of spcl_problem_f32 == the C compiler's layout of include/spcl.h."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "sz.c"
    fields = [f[0] for f in nat.ProblemF32._fields_]
    body = "".join(f'printf("%zu\\n", offsetof(spcl_problem_f32, {f}));' for f in fields)
    src.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "{ROOT}/include/spcl.h"\n'
                   f'int main(void){{printf("%zu\\n", sizeof(spcl_problem_f32));{body}return 0;}}\n')
    exe = tmp_path / "sz"
    subprocess.run([cc, str(src), "-o", str(exe)], check=True)
    out = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(nat.ProblemF32)
    assert out[1:] == [getattr(nat.ProblemF32, f).offset for f in fields]


def test_group_abi_validates_arguments_without_touching_the_gpu():
    h = nat.lib()
    assert h.spcl_supcon_group_fwd_f32(None, 3, None) == -1
    assert h.spcl_supcon_group_fwd_f32(ctypes.byref((nat.ProblemF32 * 9)()), 9, None) == -2       # > SPCL_MAX_GROUP
    assert h.spcl_supcon_group_bwd_f32(ctypes.byref((nat.ProblemF32 * 2)()), 2, None) == -1       # null operands
    import spcl_b200
    with pytest.raises(ValueError):
        spcl_b200.grouped_forward([spcl_b200.SupConLoss1()], [], None)


def test_fused_small_batch_abi_without_a_gpu(monkeypatch):
    """The one-launch entry point validates like the staged ones, reports no capacity where no device answers, and the
    host-side route selection follows the capacity and the SPCL_FUSED_SMALL switch."""
    import re
    h = nat.lib()
    assert h.spcl_supcon_group_fused_f32(None, 3, None) == nat.ERR_INVALID_ARG
    assert h.spcl_supcon_group_fused_f32(ctypes.byref((nat.ProblemF32 * 9)()), 9, None) == nat.ERR_UNSUPPORTED
    assert h.spcl_supcon_group_fused_f32(ctypes.byref((nat.ProblemF32 * 2)()), 2, None) == nat.ERR_INVALID_ARG
    if not torch.cuda.is_available():
        assert h.spcl_supcon_fused_capacity() == 0
    header = (ROOT / "include" / "spcl.h").read_text()
    codes = dict(re.findall(r"#define (SPCL_ERR_\w+) \((-\d+)\)", header))
    assert {k: int(v) for k, v in codes.items()} == {
        "SPCL_ERR_INVALID_ARG": nat.ERR_INVALID_ARG, "SPCL_ERR_UNSUPPORTED": nat.ERR_UNSUPPORTED,
        "SPCL_ERR_CUDA": nat.ERR_CUDA, "SPCL_ERR_NO_DRIVER": nat.ERR_NO_DRIVER}
    from spcl_b200 import ops
    monkeypatch.setattr(ops, "fused_capacity", lambda device: 296)
    assert ops.fused_fits([(256, 256)] * 3, "cuda")                  # cfg2: 3 x 8 x 8 tiles
    assert ops.fused_fits([(544, 96)], "cuda")                       # 17 x 17 tiles
    assert not ops.fused_fits([(545, 96)], "cuda")                   # 18 x 18
    assert not ops.fused_fits([(256, 256)] * 5, "cuda")              # 320 CTAs
    assert not ops.fused_fits([], "cuda")
    monkeypatch.setenv("SPCL_FUSED_SMALL", "0")
    assert not ops.fused_fits([(64, 128)], "cuda")


def test_dense_tail_rejects_what_is_not_on_the_hot_path():
    import spcl_b200
    with pytest.raises(NotImplementedError):
        spcl_b200.DenseProjectionTail((16, 16), pool_name="identical")
    with pytest.raises(NotImplementedError):
        spcl_b200.DenseProjectionTail((16, 16), normalize=False)
    assert spcl_b200.DenseProjectionTail((16, 16), pool_name="adaptive_max")._pool == "max"
    assert spcl_b200.DenseProjectionTail((10, 10))._spatial_size == (10, 10)


def test_point_coordinates_are_distinct_rows_and_columns_per_image():
    """infonce.py:20-22 draws n rows and n columns without replacement: no two points of an image share either."""
    import spcl_b200
    pts = spcl_b200.point_coordinates(16, 10, 12, 5, seed=9).numpy()
    assert pts.min() >= 0 and pts.max() < 120
    for row in pts:
        assert len(set(row // 12)) == 5 and len(set(row % 12)) == 5
    with pytest.raises(ValueError):
        spcl_b200.point_coordinates(1, 4, 4, 5, seed=0)            # more points than rows: numpy refuses, like the reference
