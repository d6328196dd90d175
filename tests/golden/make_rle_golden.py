"""Collects the COCO RLE strings the reference's own tests hold (real pycocotools output) into tests/golden/rle_golden.json.
Run in the build container (needs /root/reference):  python tests/golden/make_rle_golden.py"""
import json
import os
import re

REF = "/root/reference"
SOURCES = ["tests/data/test_coco_evaluation.py", "tests/test_visualizer.py"]


def main():
    cases = []
    for rel in SOURCES:
        text = open(os.path.join(REF, rel)).read()
        for m in re.finditer(r'"counts":\s*"([^"]+)"', text):
            line = text.count("\n", 0, m.start()) + 1
            window = text[max(0, m.start() - 200):m.end() + 200]
            size = re.search(r'"size":\s*\[(\d+),\s*(\d+)\]', window)
            cases.append({"source": f"{rel}:{line}", "size": [int(size.group(1)), int(size.group(2))], "counts": m.group(1)})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rle_golden.json")
    json.dump(cases, open(out, "w"), indent=1)
    print(f"{len(cases)} strings -> {out}")


if __name__ == "__main__":
    main()
