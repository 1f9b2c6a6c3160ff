"""CPU prediction of the device path's error per full-depth fixture case: the oracle forward with GEMM operands rounded to
fp16 / bf16 (oracle/vit_oracle.py `operand_dtype`) against the committed reference-generated fixture.  Used to set the GPU
test tolerances without a GPU; the GPU tests print the measured values next to these."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from law_of_vision_representation_in_mllms_b200.weights import canonical_weights   # noqa: E402
from oracle import vit_oracle                                                       # noqa: E402
from oracle.make_golden import VIT_CASES, case_pos_interp, vit_case_inputs          # noqa: E402

torch.set_num_threads(min(32, os.cpu_count() or 1))
for case in VIT_CASES:
    name, key, nl, size, n, sl, sf = case
    if nl is not None and "--all" not in sys.argv:
        continue
    rec = torch.load(os.path.join(ROOT, "tests", "golden", f"vit_{name}.pt"))
    spec, sd, imgs = vit_case_inputs(case)
    cw = canonical_weights(spec, sd, image_size=size, pos_interp=case_pos_interp(name))
    res = {}
    for od in (None, torch.float16, torch.bfloat16):
        out = vit_oracle.tower_forward(cw, spec.at_resolution(size), imgs, sl, sf, operand_dtype=od)
        got = out[:, rec["rows"]]
        res[str(od)] = ((got - rec["values"]).norm() / rec["values"].norm()).item()
    print(f"{name:28s} fp32 {res['None']:.2e}  fp16 {res['torch.float16']:.2e}  bf16 {res['torch.bfloat16']:.2e}  "
          f"ref_fp16 {rec.get('ref_fp16_rel')}  ref_bf16 {rec.get('ref_bf16_rel')}", flush=True)
