"""GPU (-m gpu): the UNMODIFIED reference modules (REF/model.py placed under baseline/_ref by oracle/place_reference.py)
run on the same B200 through stock PyTorch -- fp32 and bf16 autocast -- beside this library, same weights, same clips
(SURVEY.md section 8d "optional honesty row": the only Blackwell path that exists without this library).  Prints both
rates; asserts that the two implementations agree on the tags and that the CUDA path is the faster one."""
import os
import sys
import time

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from oracle import place_reference, ref_loader  # noqa: E402
from wfl_asr_b200 import synth  # noqa: E402
from wfl_asr_b200.model import BIOPhonemeTagger  # noqa: E402

DEV = torch.device("cuda:0")


@pytest.mark.skipif(not place_reference.placed(), reason="baseline/_ref not placed (run __graft_entry__.build() where /root/reference exists)")
def test_reference_modules_on_the_same_gpu():
    import numpy as np
    cfg = synth.workload_config("cfg2")
    labels = synth.synth_labels(30)
    ours = synth.bench_model(BIOPhonemeTagger, cfg, labels)
    sd = {k: v.detach().clone() for k, v in ours.state_dict().items()}
    ours = ours.to(DEV).eval()
    ref_loader.REF_DIR = place_reference.DEST
    ref = ref_loader.build_reference_model(cfg, labels, randomize_bn=False)
    ref.load_state_dict(sd, strict=True)
    ref = ref.to(DEV).eval()
    B = 8
    wave = torch.from_numpy(np.stack([synth.synth_wave(i, 30.0) for i in range(B)]).astype(np.float32)).to(DEV)
    lang = torch.zeros(B, dtype=torch.long, device=DEV)

    def rate(fn, reps):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        torch.cuda.synchronize()
        return B * 30.0 * reps / (time.perf_counter() - t0), out

    with torch.no_grad():
        r32, (l32, _) = rate(lambda: ref(wave, lang), 2)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            r16, (l16, _) = rate(lambda: ref(wave, lang), 2)
        rours, (lo, _) = rate(lambda: ours(wave, lang), 5)
    a_ours = (lo.argmax(-1) == l32.argmax(-1)).float().mean().item()
    a_bf16 = (l16.float().argmax(-1) == l32.argmax(-1)).float().mean().item()
    e_ours = (lo - l32).abs().max().item() / l32.abs().max().item()
    e_bf16 = (l16.float() - l32).abs().max().item() / l32.abs().max().item()
    print(f"\n[reference on B200, cfg2 model, {B} x 30 s, wall clock incl. its host-side feature extractor]"
          f"\n  REF/model.py fp32 (stock PyTorch)   : {r32:9.1f} audio-s/s"
          f"\n  REF/model.py bf16 autocast          : {r16:9.1f} audio-s/s   tags vs fp32 {a_bf16:.4%}, logits rel err {e_bf16:.2e}"
          f"\n  wfl_asr_b200 model.forward (batch {B}): {rours:9.1f} audio-s/s   tags vs fp32 {a_ours:.4%}, logits rel err {e_ours:.2e}")
    assert a_ours >= 0.999 and e_ours <= 1e-2
    assert rours > r16 and rours > r32
