"""Times the encoder_type "none" model (REF/model.py:82-91: mel-power features, d = 80; BiLSTM 2 + 2 Conformer + dilated,
random init) on a batch of 30 s clips."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wfl_asr_b200 import synth
from wfl_asr_b200.model import BIOPhonemeTagger

B = int(os.environ.get("B", "32"))
cfg = copy.deepcopy(synth.BASE_CONFIG)
cfg["model"]["encoder_type"] = "none"
torch.manual_seed(0)
model = BIOPhonemeTagger(cfg, synth.synth_labels(30)).cuda().eval()
wave = torch.rand(B, 480000) * 2 - 1
wave = (wave / wave.abs().amax(dim=1, keepdim=True)).cuda()
lang = torch.zeros(B, dtype=torch.long, device="cuda")
for _ in range(3):
    model(wave, lang)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    logits, _ = model(wave, lang)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"encoder none: batch {B} x 30 s, T {logits.shape[1]}: {ms:.3f} ms/pass = {B * 30.0 / ms * 1e3:.0f} audio-s/s")
