"""The README usage snippet as a runnable script (one B200): encoder -> flow training steps with the fused loss and
FlatAdamW, sampling, one denoising step."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, nf_b200
flow = nf_b200.RealNVP(8, 256, 64, 6, prior=None, precision="fp32").cuda()
enc = nf_b200.GEncoder(58, 64).cuda()
opt = nf_b200.FlatAdamW(list(flow.parameters()) + list(enc.parameters()), flows=[flow], lr=3e-4, weight_decay=1e-6)
obs, action = torch.randn(4096, 58, device="cuda"), torch.randn(4096, 8, device="cuda")
z, outputs, log_dets = flow(action, enc(obs))
for i in range(3):
    loss = flow.nll(action, enc(obs))
    opt.zero_grad(); loss.backward(); opt.step()
    print("loss", loss.item())
with torch.no_grad():
    y = enc(obs)
    a = flow.reverse(torch.randn(4096, 8, device="cuda"), y)
a = flow.denoise(a, y, noise_std=0.1)
print("ok", a.shape, torch.isfinite(a).all().item())
