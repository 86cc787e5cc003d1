"""CPU-only tests: host logic of the drop-in (packing plans, arena layout, range merging, sharding), the
C-ABI library (loads, exports every symbol include/sn100.h declares -- no compute calls without a GPU), and
the loud failure of the product path when there is no CUDA device."""
import os
import re
import ctypes

import pytest
import torch

import icei_b200 as sn
from icei_b200 import _lib
from icei_b200.arena import ParamArena

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "sn100.h")).read()
    declared = set(re.findall(r"\b(sn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libsn100.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert _lib.load().sn_version() == 100


def test_no_cpu_fallback():
    dec = sn.DecoderFactoredLSTM(12, 16, 16, 53, 1)
    cap = torch.randint(4, 53, (3, 5))
    with pytest.raises(RuntimeError):
        dec(cap, [5, 4, 3], torch.randn(3, 12), teacher_forcing_ratio=1.0)
    with pytest.raises(RuntimeError):
        sn.ops.linear_nt(torch.randn(4, 4), torch.randn(4, 4))


@pytest.mark.parametrize("lengths", [[7], [5, 5, 5], [9, 7, 7, 4, 1], [20] * 96, list(range(30, 0, -1))])
def test_pack_plan_matches_pack_padded_sequence(lengths):
    plan = sn.get_plan(lengths)
    B, T = len(lengths), max(lengths)
    x = torch.arange(B * T).view(B, T)
    packed = torch.nn.utils.rnn.pack_padded_sequence(x, lengths, batch_first=True)
    assert plan.bs == packed.batch_sizes.tolist()
    assert plan.N == sum(lengths) == packed.data.numel()
    got = x[torch.from_numpy(plan.row_b_np).long(), torch.from_numpy(plan.row_t_np).long()]
    assert torch.equal(got, packed.data)
    for t in range(plan.T):
        assert plan.off[t] == sum(plan.bs[:t])
    assert sn.get_plan(lengths) is plan     # cached


def test_pack_plan_rejects_bad_lengths():
    with pytest.raises(RuntimeError):
        sn.batch_sizes_from_lengths([3, 5])
    with pytest.raises(RuntimeError):
        sn.batch_sizes_from_lengths([3, 0])
    with pytest.raises(ValueError):
        sn.batch_sizes_from_lengths([])


def test_shard_lengths_sorted_and_balanced():
    lengths = sorted([20, 19, 19, 17, 15, 15, 12, 11, 9, 9, 7, 5, 5, 4, 3, 2], reverse=True)
    seen = []
    for r in range(4):
        idx, ls = sn.shard_lengths(lengths, 4, r)
        assert ls == sorted(ls, reverse=True)
        sn.batch_sizes_from_lengths(ls)
        seen += idx
    assert sorted(seen) == list(range(len(lengths)))


@pytest.mark.parametrize("make", [
    lambda: sn.DecoderFactoredLSTM(12, 16, 20, 53, 1),
    lambda: sn.DecoderRNN(12, 16, 53, 1),
    lambda: sn.DecoderFactoredLSTMAtt(16, 12, 16, 20, 53, 1, feature_size=24),
    lambda: sn.DecoderRNNAtt(16, 12, 16, 53, 1, feature_size=24),
])
def test_arena_binding_and_state_dict_interchange(make):
    torch.manual_seed(0)
    dec = make()
    ref_sd = {k: v.clone() for k, v in dec.state_dict().items()}
    a = dec.arena()
    assert a.bound()
    for n, p in dec.named_parameters():                      # values survive, storage is the arena
        assert torch.equal(p.data, ref_sd[n])
        assert p.data_ptr() == a.flat.data_ptr() + 4 * a.offset[n]
        assert a.offset[n] % 1 == 0
    # groups are contiguous stacks
    for g, (start, n) in zip(a.groups, a.group_span):
        assert start % 64 == 0
        assert sum(a.numel[x] for x in g) == n
    # load_state_dict copies in place: still bound
    dec.load_state_dict({k: v + 1 for k, v in ref_sd.items()})
    assert a.bound()
    # .double().float() replaces storages: re-bind on demand, values kept
    dec2 = make()
    dec2.load_state_dict(ref_sd)
    dec2.arena()
    dec2.double().float()
    a2 = dec2.arena()
    assert a2.bound() and a2.version == 2
    for n, p in dec2.named_parameters():
        assert torch.equal(p.data, ref_sd[n])


def test_reference_parameter_names():
    """The drop-in keeps the reference's parameter names and shapes (checkpoint interchange)."""
    from oracle import port
    pairs = [
        (sn.DecoderFactoredLSTM(12, 16, 20, 53, 1), port.DecoderFactoredLSTM(12, 16, 20, 53, 1)),
        (sn.DecoderRNN(12, 16, 53, 1), port.DecoderRNN(12, 16, 53, 1)),
        (sn.DecoderFactoredLSTMAtt(16, 12, 16, 20, 53, 1, feature_size=24),
         port.DecoderFactoredLSTMAtt(16, 12, 16, 20, 53, 1, feature_size=24)),
        (sn.DecoderRNNAtt(16, 12, 16, 53, 1, feature_size=24), port.DecoderRNNAtt(16, 12, 16, 53, 1, feature_size=24)),
    ]
    for mine, ref in pairs:
        a = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
        b = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        assert a == b
        mine.load_state_dict(ref.state_dict())
    assert len(pairs[0][0].state_dict()) == 59 and len(pairs[2][0].state_dict()) == 89


def test_merged_ranges_and_grad_ranges():
    dec = sn.DecoderFactoredLSTM(12, 16, 20, 53, 1)
    a = dec.arena()
    names = dec._seq_grad_names("happy") + list(dec._out_names())
    rs = sn.merged_ranges(a, names)
    covered = sum(n for _, n in rs)
    assert covered >= sum(a.numel[n] for n in names)
    # inactive styles are not covered
    for o, n in rs:
        for s in ("S_ff", "S_sad_i", "S_angry_c"):
            off = a.offset[s + ".weight"]
            assert not (o <= off < o + n), s
    # publish + grad_ranges round trip
    a.publish_grads(names, a.gflat)
    items, foreign = a.grad_ranges()
    assert not foreign and {n for _, _, n in items} == set(names)
    assert dec.S_sad_i.weight.grad is None and dec.S_happy_i.weight.grad is not None
    dec.S_happy_i.weight.grad = torch.zeros_like(dec.S_happy_i.weight)   # a foreign gradient tensor
    items, foreign = a.grad_ranges()
    assert foreign == ["S_happy_i.weight"]


def test_stack_module_layout_and_state_dict_interchange():
    """configs[3] stack: parameter count of the survey (73 395 664 at E=300, H=512, F=1024, V=10000, 3 layers), reference
    names for layer 0, ``l{l}_`` names above, one arena, per-layer state_dict loading from oracle-port decoders."""
    from oracle import port
    dec = sn.DecoderFactoredLSTMStack(300, 512, 1024, 10000, 3)
    assert sum(p.numel() for p in dec.parameters()) == 73395664
    ref_names = set(dict(sn.DecoderFactoredLSTM(12, 16, 20, 53, 1).named_parameters()))
    small = sn.DecoderFactoredLSTMStack(12, 16, 20, 53, 3)
    names = set(dict(small.named_parameters()))
    assert ref_names <= names
    upper = {n for n in names if n.startswith("l1_") or n.startswith("l2_")}
    assert names == ref_names | upper
    assert {n[3:] for n in upper if n.startswith("l1_")} == {n for n in ref_names if n[0] in "UVWS"}
    assert dict(small.named_parameters())["l1_V_i.weight"].shape == (20, 16)      # embed_size = hidden_size above layer 0
    a = small.arena()                                                             # binds on CPU tensors too
    assert a.total >= sum(p.numel() for p in small.parameters())
    H, F = 16, 20
    blk = a.block(["l2_" + "W_" + g + ".weight" for g in "ifoc"], (4 * H, H))    # gate-stacked and contiguous
    assert blk.data_ptr() == small.l2_W_i.weight.data_ptr()
    layers = [port.DecoderFactoredLSTM(12 if l == 0 else 16, 16, 20, 53, 1) for l in range(3)]
    small.load_layer_state_dicts([m.state_dict() for m in layers])
    assert torch.equal(small.B.weight, layers[0].B.weight) and torch.equal(small.C.weight, layers[0].C.weight)
    assert torch.equal(small.l2_S_happy_o.weight, layers[2].S_happy_o.weight)
    g = small._seq_grad_names("sad")
    assert "l1_S_sad_i.weight" in g and "l1_S_happy_i.weight" not in g and "l2_W_c.bias" in g and "C.weight" not in g
    with pytest.raises(ValueError):
        small.load_layer_state_dicts([layers[0].state_dict()])


def test_multitask_schedule_order_and_modes():
    """factual pass on the first trainer, then one emotion pass per available tag on the second
    (stylenet/train_multitask.py:192-235)."""
    calls = []

    class FakeTrainer:
        def __init__(self, name):
            self.name = name

        def step(self, cap, lengths, feat, mode):
            calls.append((self.name, mode, cap))
            return (len(calls),), None

    sched = sn.MultitaskSchedule(FakeTrainer("A"), FakeTrainer("B"), tags=("happy", "sad", "angry"))
    out = sched.step(("cf", [3], "ff"), {"sad": ("cs", [3], "fs"), "happy": ("ch", [3], "fh")})
    assert calls == [("A", "factual", "cf"), ("B", "happy", "ch"), ("B", "sad", "cs")]
    assert list(out) == ["factual", "happy", "sad"]


def test_split_k_heuristic_and_host_step_tables():
    from icei_b200 import ops
    assert ops._pick_splits(1920, 512, 10000, 1) > 1          # dH = dL C: few tiles, K = vocabulary
    assert ops._pick_splits(2048, 512, 1920, 1) == 1          # weight gradient: 30 k-blocks, unsplit
    assert ops._pick_splits(1920, 10000, 512, 1) == 1         # plenty of tiles
    plan = sn.get_plan([5, 3, 3, 1])
    bs, off = ops._host_steps(plan)
    assert list(bs) == plan.bs and list(off) == plan.off[:plan.T]
    assert ops._host_steps(plan) is plan.__dict__["_host_steps"]


def test_header_and_ctypes_signatures_agree_on_arity():
    """Every declaration in include/sn100.h has as many parameters as its ctypes signature in _lib.py (a silent
    mismatch would shift every argument after it)."""
    hdr = open(os.path.join(ROOT, "include", "sn100.h")).read()
    seen = 0
    for m in re.finditer(r"\b(?:int32_t|int64_t|const char\*)\s+(sn_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, re.S):
        name, args = m.group(1), m.group(2)
        n = 0 if args.strip() in ("void", "") else len(args.split(","))
        assert name in _lib.SIGNATURES, name
        assert len(_lib.SIGNATURES[name][1]) == n, (name, n, len(_lib.SIGNATURES[name][1]))
        seen += 1
    assert seen == len(_lib.SIGNATURES)


def test_arena_rebinds_when_a_parameter_object_is_replaced():
    """``decoder.B.weight = nn.Parameter(pretrained)`` (a new Parameter OBJECT in the middle of the layout) must be
    picked up: the kernels read the arena, so a stale snapshot would silently keep using the old values."""
    torch.manual_seed(0)
    dec = sn.DecoderFactoredLSTM(12, 16, 20, 53, 1)
    a = dec.arena()
    v0 = a.version
    new_w = torch.nn.Parameter(torch.full((16, 20), 0.25))
    dec.U_f.weight = new_w                      # neither the first nor the last parameter of the layout
    a2 = dec.arena()
    assert a2.version == v0 + 1
    assert a2.named["U_f.weight"] is dec.U_f.weight
    o, n = a2.offset["U_f.weight"], a2.numel["U_f.weight"]
    assert torch.equal(a2.flat[o:o + n], torch.full((n,), 0.25))
    assert dec.U_f.weight.data_ptr() == a2.flat.data_ptr() + 4 * o
    # unrelated registrations elsewhere only cost one identity walk, no rebuild
    torch.nn.Linear(3, 3)
    assert dec.arena().version == v0 + 1


def test_decoder_and_optimizer_pickle_round_trip():
    """save_checkpoint pickles the whole decoder and optimizers (stylenet/utils.py:62-90); per-process caches (CUDA
    graphs, streams, events) must not travel, parameters and Adam state must."""
    import io
    import threading
    torch.manual_seed(0)
    dec = sn.DecoderFactoredLSTM(12, 16, 20, 53, 1)
    dec.arena()
    opt = sn.FusedClampAdam(dec, lr=3e-4)
    a = opt._state()
    opt.m.uniform_(-1, 1); opt.v.uniform_(0, 1); opt.steps_dev.fill_(3)
    # what a greedy forward / sample() leave behind: objects that cannot be pickled
    dec.__dict__["_greedy_graphs"] = {"k": threading.Lock()}
    dec.__dict__["_decode_sessions"] = {"k": threading.Lock()}
    dec.__dict__["_side_streams"] = {0: threading.Lock()}
    dec.__dict__["_out_w16_ev"] = threading.Lock()
    buf = io.BytesIO()
    torch.save({"decoder": dec, "optimizer": opt}, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    dec2, opt2 = back["decoder"], back["optimizer"]
    assert "_greedy_graphs" not in dec2.__dict__ and "_arena" not in dec2.__dict__
    for (n, p), (n2, q) in zip(dec.state_dict().items(), dec2.state_dict().items()):
        assert n == n2 and torch.equal(p, q)
    assert opt2.decoder is dec2
    a2 = opt2._state()                          # re-binds to the un-pickled decoder's new arena, keeps the moments
    assert torch.equal(opt2.m, opt.m) and torch.equal(opt2.v, opt.v)
    assert opt2.steps_dev.tolist() == opt.steps_dev.tolist()
    assert a2.total == a.total


def test_optimizer_state_dict_uses_the_torch_adam_schema():
    torch.manual_seed(0)
    dec = sn.DecoderFactoredLSTM(12, 16, 20, 53, 1)
    opt = sn.FusedClampAdam(dec, lr=3e-4)
    a = opt._state()
    opt.m.uniform_(-1, 1); opt.v.uniform_(0, 1)
    stepped = ["C.weight", "C.bias", "S_happy_i.weight"]
    for n in stepped:
        opt.steps_dev[opt.index[n]] = 7
    sd = opt.state_dict()
    params = opt.param_groups[0]["params"]
    names = {id(p): n for n, p in dec.named_parameters()}
    assert sorted(names[id(params[i])] for i in sd["state"]) == sorted(stepped)
    g = sd["param_groups"][0]
    assert g["lr"] == 3e-4 and g["betas"] == (0.9, 0.999) and g["eps"] == 1e-8 and g["params"] == list(range(len(params)))
    for i, st in sd["state"].items():
        n = names[id(params[i])]
        assert float(st["step"]) == 7.0 and st["exp_avg"].shape == params[i].shape
        assert torch.equal(st["exp_avg_sq"].reshape(-1), opt.v[a.offset[n]:a.offset[n] + a.numel[n]])
    # torch.optim.Adam accepts it (same schema) ...
    ref = torch.optim.Adam(dec.parameters(), lr=1.0)
    ref.load_state_dict(sd)
    assert ref.param_groups[0]["lr"] == 3e-4
    # ... and a fresh FusedClampAdam restores moments, step counters and lr from it
    dec2 = sn.DecoderFactoredLSTM(12, 16, 20, 53, 1)
    opt2 = sn.FusedClampAdam(dec2, lr=1.0)
    opt2.load_state_dict(sd)
    a2 = opt2._state()
    for n in stepped:
        o, k = a2.offset[n], a2.numel[n]
        assert torch.equal(opt2.m[o:o + k], opt.m[o:o + k]) and opt2.steps_dev[opt2.index[n]].item() == 7
    o, k = a2.offset["V_i.weight"], a2.numel["V_i.weight"]
    assert not opt2.m[o:o + k].any()             # parameters without state start from zero moments
    assert opt2.param_groups[0]["lr"] == 3e-4


def test_encoder_signatures_and_state_dict_names_match_reference():
    import inspect
    import torch.nn as nn
    assert list(inspect.signature(sn.EncoderCNN.__init__).parameters)[:2] == ["self", "embed_size"]
    assert list(inspect.signature(sn.EncoderCNNAtt.__init__).parameters)[:2] == ["self", "encoded_image_size"]
    assert inspect.signature(sn.EncoderCNNAtt.__init__).parameters["encoded_image_size"].default == 14
    trunk = nn.Sequential(nn.Conv2d(3, 8, 3), nn.AdaptiveAvgPool2d((1, 1)))
    enc = sn.EncoderCNN(12, backbone=trunk, in_features=8)
    keys = list(enc.state_dict())
    assert keys[:2] == ["resnet.0.weight", "resnet.0.bias"]
    assert keys[2:] == ["linear.weight", "linear.bias", "bn.weight", "bn.bias", "bn.running_mean", "bn.running_var",
                        "bn.num_batches_tracked"]
    assert enc.bn.momentum == 0.01
    with pytest.raises(RuntimeError):
        enc(torch.randn(2, 3, 8, 8))            # CPU tensors: the tail has no CPU fallback
    att = sn.EncoderCNNAtt(backbone=nn.Sequential(nn.Conv2d(3, 8, 3)))
    assert att.encoded_image_size == 14 and list(att.state_dict()) == ["resnet.0.weight", "resnet.0.bias"]


def test_dp_receive_slots_and_range_merging():
    """Host side of the push-form exchange: a rank's receive slot holds its share of the arena's 4096-element chunks
    (sn_dp_slot_elems is plain host arithmetic: callable without a GPU); ranges of one bucket are merged when they touch."""
    from icei_b200 import ops
    for elems, world in [(1, 1), (4096, 2), (4097, 2), (12_400_000, 8), (73_400_000, 8), (5, 3)]:
        chunks = -(-elems // 4096)
        want = -(-chunks // world) * 4096
        assert ops.dp_slot_elems(elems, world) == want
        # every owned chunk `aid` (aid % world == rank) has a place: slot index aid // world < slot chunks
        assert (chunks - 1) // world < want // 4096
    assert ops.dp_slot_elems(-1, 2) == -1 and ops.dp_slot_elems(10, 0) == -1
    assert ops.merge_adjacent([(100, 20), (0, 50), (50, 50), (130, 5)]) == [(0, 120), (130, 5)]
    assert ops.merge_adjacent([]) == []


def test_arena_content_key_sees_every_kind_of_weight_change():
    """The few-row decode caches the collapsed U S V chain on ParamArena.content_key(): it must change on in-place torch
    updates of a Parameter, on load_state_dict, on a re-bind, and when the optimizer kernels / graph replays bump
    kernel_epoch -- and stay put otherwise."""
    import icei_b200 as sn
    torch.manual_seed(0)
    dec = sn.DecoderFactoredLSTM(12, 16, 8, 30, 1, dropout=0.0)
    a = dec.arena()
    k0 = a.content_key()
    assert a.content_key() == k0
    with torch.no_grad():
        dec.U_i.weight.mul_(2.0)
    k1 = dec.arena().content_key()
    assert k1 != k0
    dec.load_state_dict({k: v.clone() for k, v in dec.state_dict().items()})
    k2 = dec.arena().content_key()
    assert k2 != k1
    dec.arena().kernel_epoch += 1
    k3 = dec.arena().content_key()
    assert k3 != k2
    dec.C.weight = torch.nn.Parameter(dec.C.weight.detach().clone())
    assert dec.arena().content_key() != k3
