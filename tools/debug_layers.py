"""Layer-by-layer comparison of the engine workspace against the bf16-emulating oracle (debug aid)."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hyperpri_oracle as O
from hyperpri_b200.src.Experiments.models import UNet

def nchw(t): return t.float().permute(0, 3, 1, 2).cpu()
def err(a, b): return ((a - b).abs().max() / b.abs().max()).item(), ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()

n, h, w = 2, 96, 136
net = UNet(3, 1, bilinear=False)
sd = O.synth_state_dict(O.unet_schema(3, 1, "unet"), 0)
net.load_state_dict(sd); net = net.cuda().train()
x = O.synth_cube(0, n, 3, h, w)
with torch.no_grad():
    net.train()
    eng = net._get_engine(torch.device("cuda", 0))
    eng.forward(x.cuda(), True)
torch.cuda.synchronize()
ws = eng.ws
O.emulate_bf16_storage(True)
q = O.q
def cbr(t, pre, i, quant=True):
    raw = q(F.conv2d(t, q(sd[f"{pre}.{i}.weight"]), None, padding=1))
    y = torch.relu(O.batch_norm(raw + sd[f"{pre}.{i}.bias"].view(1, -1, 1, 1), sd, f"{pre}.{i+1}", True, None))
    return raw, (q(y) if quant else y)
with torch.no_grad():
    t = q(x)
    print("x", err(nchw(ws["x"])[:, :3], t))
    acts = []
    for l in range(5):
        pre = "inc.double_conv" if l == 0 else f"down{l}.maxpool_conv.1.double_conv"
        raw, a = cbr(t, pre, 0)
        print(l, "raw_a", err(nchw(ws[f"enc_raw_a{l}"]), raw), "act_a", err(nchw(ws[f"enc_act_a{l}"]), a))
        raw, b = cbr(a, pre, 3, quant=False)
        print(l, "raw_b", err(nchw(ws[f"enc_raw_b{l}"]), raw))
        if l < 4:
            C = raw.shape[1]
            print(l, "skip", err(nchw(ws[f"cat{l}"][..., :C]), q(b)))
            t = q(F.max_pool2d(b, 2))
            print(l, "pool", err(nchw(ws[f"pool{l+1}"]), t))
            acts.append(b)
        else:
            t = q(b)
            print(l, "act_b4", err(nchw(ws["act_b4"]), t))
    for l in (3, 2, 1, 0):
        i = 4 - l
        up = q(F.conv_transpose2d(t, q(sd[f"up{i}.up.weight"]), sd[f"up{i}.up.bias"], stride=2))
        sk = acts[l]
        dy, dx = sk.shape[2] - up.shape[2], sk.shape[3] - up.shape[3]
        up = F.pad(up, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        C = sk.shape[1]
        print(l, "up", err(nchw(ws[f"cat{l}"][..., C:]), up))
        cat = torch.cat([q(sk), up], 1)
        pre = f"up{i}.conv.double_conv"
        raw, a = cbr(cat, pre, 0)
        print(l, "dec_raw_a", err(nchw(ws[f"dec_raw_a{l}"]), raw), "act", err(nchw(ws[f"dec_act_a{l}"]), a))
        raw, b = cbr(a, pre, 3, quant=(l > 0))
        print(l, "dec_raw_b", err(nchw(ws[f"dec_raw_b{l}"]), raw))
        t = b
    lg = F.conv2d(t, sd["outc.conv.weight"], sd["outc.conv.bias"])
    print("logits", err(ws["logits"].cpu(), lg))
