"""Generates tests/golden/reference_facts.json by importing the parts of the reference that DO import in this
container (the BPE tokenizer and the YAML configs; TensorFlow 1.15 itself is not installable -- SURVEY.md 8(c)).
Run here (needs /root/reference):  python tests/golden/make_golden.py"""
import json
import os
import sys

import yaml

REF = "/root/reference"
sys.path.insert(0, REF)
from utils.encode.encoder import MASK, PADDING, START, get_encoder  # noqa: E402

enc = get_encoder()
texts = [" answer question:", "Hello world", " the quick brown fox", "MERLOT reserve"]
facts = {
    "tokenizer": {"len": len(enc.encoder), "max_id": max(enc.encoder.values()) + 100 if False else None,
                  "specials": {"PADDING": PADDING, "MASK": MASK, "START": START},
                  "encode": {t: enc.encode(t) for t in texts}},
    "configs": {},
}
for name in ("merlot.yaml", "merlot_5segments.yaml"):
    with open(os.path.join(REF, "model/configs", name)) as f:
        c = yaml.load(f, Loader=yaml.FullLoader)
    facts["configs"][name] = {"model": c["model"], "optimizer": c["optimizer"], "data": {k: c["data"][k] for k in
                              ("num_chunks", "chunk_text_len") if k in c["data"]}}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_facts.json")
json.dump(facts, open(out, "w"), indent=1, sort_keys=True)
print("wrote", out)
