import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from premvos_b200 import pwc, synth, _lib
sd = synth.pwc_synthetic_state_dict(0)
net = pwc.pwc_dc_net(None); net.load_state_dict(sd); net.cuda().eval()
x = torch.from_numpy(synth.synthetic_pwc_input(4, 448, 1024, seed=1)).cuda()
for _ in range(3): net(x)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): net(x)
e1.record(); torch.cuda.synchronize()
_lib.profile_begin(); net(x); torch.cuda.synchronize(); p=_lib.profile_end()
print(os.environ.get("PREMVOS_CORR_TMA_MIN_PX"), os.environ.get("PREMVOS_CORR_TMA"), "pwc batch 4: %.3f ms per forward; corr %s" % (e0.elapsed_time(e1)/20, p["corr81_cp8_kernel"]))
