"""Generates the golden fixtures under tests/golden/ (run in the BUILD container only).

Sources of truth:
  * OpenJPEG 2.5.4 (the library go-dicom-codec's JPEG 2000 path clones arithmetic-for-arithmetic,
    jpeg2000/encoder.go:1790) reached through Pillow 12.2.0's bundled libopenjp2;
  * the reference repository's own raw fixtures (test-data/htj2k/interop/*/input.raw, test-data/CT1_J2KI).
Neither exists on the GPU box, so the outputs are committed and this script is kept as their provenance.

    python tests/golden/make_golden.py [/root/reference]
"""
import io
import json
import os
import shutil
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def smooth_noise(rng, h, w, bits, sigma_div=256):
    mid, amp = 2 ** (bits - 1), 2 ** (bits - 2)
    yy, xx = np.mgrid[0:h, 0:w]
    v = mid + amp * np.sin(xx / 17.0) * np.cos(yy / 23.0) + rng.normal(0, 2 ** bits / sigma_div, (h, w))
    return np.clip(np.rint(v), 0, 2 ** bits - 1)


def qcd_bytes(cs):
    i = cs.find(b"\xff\x5c")
    ln = int.from_bytes(cs[i + 2:i + 4], "big")
    return cs[i + 4:i + 2 + ln]  # Sqcd + SPqcd


def encode(arr, levels, irreversible, mct=0):
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="JPEG2000", irreversible=irreversible, num_resolutions=levels + 1,
                              no_jp2=True, mct=mct)
    return buf.getvalue()


def decode(cs, reduce=0):
    im = Image.open(io.BytesIO(cs))
    im.reduce = reduce
    im.load()
    return np.asarray(im)


def main():
    import PIL
    meta = {"pillow": PIL.__version__, "openjpeg": "2.5.4 (pillow.libs libopenjp2)", "cases": {}}
    out = {}
    # ---- 9/7 + quantization + dequantization, mono: decoded pixels of OpenJPEG with all passes kept
    mono97 = [("m97_64x64_8b_L1", 64, 64, 8, 1, "noise"), ("m97_64x64_8b_L3", 64, 64, 8, 3, "smooth"),
              ("m97_67x53_8b_L2", 53, 67, 8, 2, "noise"), ("m97_96x80_16b_L2", 80, 96, 16, 2, "smooth"),
              ("m97_128x128_16b_L5", 128, 128, 16, 5, "smooth"), ("m97_101x77_16b_L3", 77, 101, 16, 3, "noise"),
              ("m97_256x256_12in16_L5", 256, 256, 12, 5, "smooth"), ("m97_333x211_12in16_L6", 211, 333, 12, 6, "noise")]
    for seed, (name, h, w, bits, L, kind) in enumerate(mono97):
        rng = np.random.default_rng(100 + seed)
        a = smooth_noise(rng, h, w, bits) if kind == "smooth" else rng.integers(0, 2 ** bits, (h, w))
        prec = 8 if bits <= 8 else 16  # Pillow has no 12-bit mode: 12-bit-range data at precision 16
        a = a.astype(np.uint8 if prec == 8 else np.uint16)
        cs = encode(a, L, True)
        out[name + "_in"] = a
        out[name + "_dec"] = decode(cs)
        out[name + "_qcd"] = np.frombuffer(qcd_bytes(cs), np.uint8)
        meta["cases"][name] = {"kind": "mono97", "w": w, "h": h, "prec": prec, "levels": L}
    # ---- RGB ICT + 9/7 (OpenJPEG's own float32 decode tail; the Go tail is deliberately different)
    for seed, (name, h, w, L) in enumerate([("c97_64x64_L3", 64, 64, 3), ("c97_130x97_L4", 97, 130, 4)]):
        rng = np.random.default_rng(200 + seed)
        a = np.stack([smooth_noise(rng, h, w, 8, 64) for _ in range(3)], -1).astype(np.uint8)
        cs = encode(a, L, True, mct=1)
        out[name + "_in"] = a
        out[name + "_dec"] = decode(cs)
        out[name + "_qcd"] = np.frombuffer(qcd_bytes(cs), np.uint8)
        meta["cases"][name] = {"kind": "rgb97", "w": w, "h": h, "prec": 8, "levels": L}
    # ---- 5/3: LL_n bands as OpenJPEG reconstructs them at reduced resolution
    for seed, (name, h, w, bits, L) in enumerate([("m53_64x64_8b_L3", 64, 64, 8, 3), ("m53_127x129_8b_L4", 129, 127, 8, 4),
                                                  ("m53_200x96_16b_L5", 96, 200, 16, 5)]):
        rng = np.random.default_rng(300 + seed)
        a = smooth_noise(rng, h, w, bits, 32).astype(np.uint8 if bits <= 8 else np.uint16)
        cs = encode(a, L, False)
        out[name + "_in"] = a
        assert np.array_equal(decode(cs), a)
        got = []
        for n in range(1, L + 1):
            try:  # Pillow's own size bookkeeping rejects some odd-size reductions; keep what decodes
                out[f"{name}_ll{n}"] = decode(cs, n)
                got.append(n)
            except OSError:
                pass
        meta["cases"][name] = {"kind": "mono53", "w": w, "h": h, "prec": bits, "levels": L, "ll_levels": got}
    # ---- RGB RCT + 5/3 LL bands
    rng = np.random.default_rng(400)
    a = np.stack([smooth_noise(rng, 72, 88, 8, 32) for _ in range(3)], -1).astype(np.uint8)
    cs = encode(a, 3, False, mct=1)
    out["c53_88x72_L3_in"] = a
    assert np.array_equal(decode(cs), a)
    meta["cases"]["c53_88x72_L3"] = {"kind": "rgb53", "w": 88, "h": 72, "prec": 8, "levels": 3}
    np.savez_compressed(os.path.join(HERE, "openjpeg_254.npz"), **out)
    # ---- the reference's raw fixtures
    idir = os.path.join(HERE, "interop")
    os.makedirs(idir, exist_ok=True)
    man = json.load(open(os.path.join(REF, "test-data/htj2k/interop/manifest.json")))
    keep = []
    for fx in man["fixtures"]:
        shutil.copyfile(os.path.join(REF, "test-data/htj2k/interop", fx["inputRaw"]), os.path.join(idir, fx["name"] + ".raw"))
        os.chmod(os.path.join(idir, fx["name"] + ".raw"), 0o644)
        keep.append({k: fx[k] for k in ("name", "width", "height", "components", "bitsAllocated", "bitsStored", "signed")})
        # every reference codestream of the fixture decodes (OpenJPEG) to input.raw: checked here, once
        for csinfo in fx["codestreams"].values():
            raw = open(os.path.join(REF, "test-data/htj2k/interop", csinfo["path"]), "rb").read()
            eoc = raw.find(b"\xff\xd9")
            try:
                d = decode(raw[:eoc + 2] if eoc >= 0 else raw)
                src = np.fromfile(os.path.join(idir, fx["name"] + ".raw"), np.uint8 if fx["bitsAllocated"] == 8 else "<u2")
                dv = d.reshape(-1).astype(np.int64)
                if fx["signed"]:  # Pillow presents signed samples offset-binary
                    dv = dv - (1 << (fx["bitsStored"] - 1))
                ok = bool(np.array_equal(dv & (0xFFFF if fx["bitsAllocated"] > 8 else 0xFF), src.astype(np.int64)))
            except Exception as e:  # HTJ2K decode support depends on the OpenJPEG build
                ok = f"not decodable here: {type(e).__name__}"
            keep[-1].setdefault("codestream_decodes_to_input", []).append(ok)
    json.dump({"source": man["source"], "fixtures": keep}, open(os.path.join(idir, "manifest.json"), "w"), indent=1)
    # ---- CT1_J2KI (the C1 image): DICOM-wrapped J2K codestream, TS .91, 512x512 signed 16-bit
    raw = open(os.path.join(REF, "test-data/CT1_J2KI"), "rb").read()
    i, j = raw.find(b"\xff\x4f\xff\x51"), raw.rfind(b"\xff\xd9")
    ct = decode(raw[i:j + 2])  # offset-binary uint16 as Pillow presents signed data
    np.savez_compressed(os.path.join(HERE, "ct1_j2ki.npz"), offset_binary=ct.astype(np.uint16))
    meta["ct1"] = {"shape": list(ct.shape), "min": int(ct.min()), "max": int(ct.max())}
    json.dump(meta, open(os.path.join(HERE, "openjpeg_254.json"), "w"), indent=1)
    print(json.dumps(meta)[:400])


if __name__ == "__main__":
    main()
