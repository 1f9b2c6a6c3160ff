"""ORACLE (test infrastructure).  Recipe that puts the reference's OWN implementation of the hot path where it can travel.

The reference is Python: nothing to compile.  Its tower wrappers import only torch + transformers, the A-score script and the
correspondence utilities only torch/numpy/PIL, so the handful of files below are COPIED from /root/reference (read-only,
present in the build container only) into oracle/_ref/ — git-ignored (reference sources never enter this repository's
history) but shipped to the GPU box with the working tree, like a compiled reference binary would be.  There
`bench.py --impl reference` and the `cpu_baseline` leg run them (`cpu_baseline.kind == "reference"`); when oracle/_ref is
absent they fall back to the restatement (`kind == "port"`).

    python oracle/build_ref.py        # also called by __graft_entry__.build() when /root/reference exists
"""
import os
import shutil

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = [
    "llava/model/multimodal_encoder/clip_encoder.py",      # CLIPVisionTower        
    "llava/model/multimodal_encoder/dinov2_encoder.py",    # DinoV2VisionTower      (a4)
    "llava/model/multimodal_encoder/siglip_encoder.py",    # SigLipVisionTower      (a5)
    "A_score/compute.py",                                  # A-score script         (a8-a10)
    "C_score/utils/utils_correspondence.py",               # correspondence / flow  (a13-a15)
]


def build_ref():
    if not os.path.isdir(REF):
        return None
    os.makedirs(OUT, exist_ok=True)
    for rel in FILES:
        shutil.copyfile(os.path.join(REF, rel), os.path.join(OUT, os.path.basename(rel)))
    with open(os.path.join(OUT, "PROVENANCE.txt"), "w") as f:
        f.write("verbatim copies from /root/reference (bronyayang/Law_of_Vision_Representation_in_MLLMs), made by "
                "oracle/build_ref.py; git-ignored; used only as the timed CPU reference arm / checker\n" + "\n".join(FILES) + "\n")
    return OUT


if __name__ == "__main__":
    print(build_ref())
