import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, torch.nn.functional as F
from oracle import disc_ref, weights
from gpu_util import rel_err
from models.domain_shift.adversarial.model import DomainDiscriminator, TinyDomainDiscriminator
for tiny in (True, False):
  for (n,h,w) in ((2,90,160),(3,128,256)):
    if not tiny and min(h,w)<64: continue
    sd = weights.discriminator_state(5, tiny=tiny)
    g = torch.Generator().manual_seed(77)
    logits = torch.randn(n, 19, h, w, generator=g) * 2
    xr = logits.clone().requires_grad_(True)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss_r, p_r = disc_ref.adversarial_bce(xr, sdr, 1.0, 0.37)
    loss_r.backward()
    for prec in ("fp32","bf16","bf16_simt"):
        m = (TinyDomainDiscriminator if tiny else DomainDiscriminator)(19)
        m.load_state_dict(sd); m.rtsds_precision = prec; m = m.cuda().train()
        x = logits.cuda().requires_grad_(True)
        p = m(F.softmax(x, dim=1))
        loss = 0.37 * F.binary_cross_entropy_with_logits(p, torch.ones_like(p))
        loss.backward()
        print(tiny,(n,h,w),prec,"out",rel_err(p.detach().cpu(), p_r.detach()),"dx",rel_err(x.grad.cpu(), xr.grad),
              {k: round(rel_err(prm.grad.cpu(), sdr[k].grad),5) for k,prm in m.named_parameters()})
