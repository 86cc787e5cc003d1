import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import icei_b200 as sn
from oracle import port
from golden_util import rel_l2
def run(dims, prec):
    A, E, H, F, V, D, S, B, T = dims
    torch.manual_seed(2)
    torch.set_default_dtype(torch.float64)
    ref = port.DecoderFactoredLSTMAtt(A, E, H, F, V, 1, feature_size=D, dropout=0.0)
    torch.set_default_dtype(torch.float32)
    dec = sn.DecoderFactoredLSTMAtt(A, E, H, F, V, 1, feature_size=D, dropout=0.0)
    dec.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    dec = dec.cuda().train().set_precision(prec)
    cap, lens, feats = port.synthetic_batch(B, T, V, feat_shape=(S, S, D), ragged=True, seed=6)
    l1 = [l - 1 for l in lens]
    tgt = port.pack_targets(cap[:, 1:], l1)
    out_ref, al_ref = ref(cap[:, :-1], l1, feats.double(), teacher_forcing_ratio=1.0, mode="angry")
    loss_ref = port.caption_loss(out_ref, tgt, al_ref)
    ref.zero_grad(); loss_ref.backward()
    out, al = dec(cap[:, :-1].cuda(), l1, feats.cuda(), teacher_forcing_ratio=1.0, mode="angry")
    loss = port.caption_loss(out, tgt.cuda(), al)
    dec.zero_grad(); loss.backward()
    print(dims, prec, 'logits', rel_l2(out.detach().cpu(), out_ref.detach()), 'alphas', rel_l2(al.detach().cpu(), al_ref.detach()), 'loss', loss.item(), loss_ref.item())
    gref = {n: p.grad for n, p in ref.named_parameters()}
    errs = {n: rel_l2(p.grad.cpu(), gref[n]) for n, p in dec.named_parameters() if gref[n] is not None and not n.endswith('full_att.bias')}
    for n, e in sorted(errs.items(), key=lambda kv: -kv[1])[:12]:
        print('   %-32s %.4f  |g|=%.3e' % (n, e, gref[n].norm().item()))
    print('   median', sorted(errs.values())[len(errs)//2])
run((64, 44, 64, 72, 600, 128, 3, 20, 9), 'bf16')
run((512, 300, 512, 512, 2000, 2048, 7, 24, 12), 'bf16')
