"""Prints the cta_group::2 stage-2 timeline of pair 0 (PPT_STAGE2_PAIR=1 PPT_PAIR_TRACE=1); debugging aid."""
import os
import sys

os.environ["PPT_STAGE2_PAIR"] = "1"
os.environ["PPT_PAIR_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from oracle import torch_port  # noqa: E402
from ppt_b200 import encoder_pack, ops  # noqa: E402

sd = torch_port.make_encoder_state()
blob = encoder_pack.pack_encoder(sd, 0).cuda()
nb = (torch.rand(128, 512, 32, 3, device="cuda") - 0.5) * 0.4
ops.encoder_forward(nb, blob, mode=0)
torch.cuda.synchronize()
os.environ["PPT_PAIR_TRACE"] = "0"
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(3):
    ops.encoder_forward(nb, blob, mode=0)
evs = []
ops.encoder_forward(nb, blob, mode=0, phase_events=evs)
torch.cuda.synchronize()
print("PHASES", {n: round(a.elapsed_time(b), 4) for n, a, b in evs}, file=sys.stderr)
