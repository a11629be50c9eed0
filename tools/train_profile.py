"""Launch-level timing of one tokenizer step with the Encoder in train mode (row f3).
    ncu --metrics gpu__time_duration.sum --csv --log-file gpurun_out/train_launches.csv python tools/train_profile.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from oracle import torch_port  # noqa: E402
from ppt_b200.tokenizer import PointTokenizer  # noqa: E402

dev = torch.device("cuda", 0)
tok = PointTokenizer(512, 32).to(dev).load_reference_state(torch_port.make_encoder_state())
tok.start_idx = 0
tok.encoder.train()
xyz = torch.rand(128, 8192, 3, device=dev) * 2 - 1
for _ in range(3):
    tok(xyz)
torch.cuda.synchronize()
