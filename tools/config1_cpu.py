#!/usr/bin/env python
"""BASELINE.json config 1: "BTS-NYU DenseNet-161 inference, 1 synthetic 416x544 image, random-init weights, on CPU
(reference tf.keras path)".

TensorFlow cannot be installed in this image, so the run is labelled **restated-reference CPU (torch)**:
  * encoder: torchvision densenet161(weights=None); taps relu0 (/2, 96 ch), pool0 (/4, 96), transition1 (/8, 192),
    transition2 (/16, 384), norm5 + ReLU (/32, 2208) -- the tensors bts.py:72 / bts_densenet.py:93-110 feed the decoder;
  * decoder: the UNMODIFIED /root/reference/bts_decoder.py (with its custom_layers.py) executed over oracle/tf_shim, the
    torch-CPU stand-in for the tf symbols the reference imports (the same arrangement that generates tests/golden/).
Build container only (needs /root/reference).  Prints one JSON line; profiles/r02_config1_cpu.json keeps the run.

    python tools/config1_cpu.py [--reps 5] [--height 416] [--width 544]
"""
import argparse
import json
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("BTS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)


def encoder_taps(model, image):
    f = model.features
    x = f.relu0(f.norm0(f.conv0(image)))
    skip_2 = x
    x = f.pool0(x)
    skip_4 = x
    x = f.transition1(f.denseblock1(x))
    skip_8 = x
    x = f.transition2(f.denseblock2(x))
    skip_16 = x
    x = f.transition3(f.denseblock3(x))
    x = torch.relu(f.norm5(f.denseblock4(x)))
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()          # noqa: E731  the decoder is NHWC (Keras channels_last)
    return [nhwc(x), nhwc(skip_2), nhwc(skip_4), nhwc(skip_8), nhwc(skip_16)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--height", type=int, default=416)
    ap.add_argument("--width", type=int, default=544)
    a = ap.parse_args()
    import torchvision
    import bts_decoder                                              # the reference file, unmodified
    from tensorflow.keras import layers as shim_layers
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    enc = torchvision.models.densenet161(weights=None).eval()
    image = torch.rand(1, 3, a.height, a.width)
    ts_enc, ts_dec = [], []
    depth = None
    with torch.no_grad():
        for rep in range(a.reps + 1):
            t0 = time.perf_counter()
            feats = encoder_taps(enc, image)
            t1 = time.perf_counter()
            shim_layers.reset(seed=1234, dtype=torch.float32)       # same random-init decoder every pass
            depth = bts_decoder.decoder_model(feats, 10.0, num_filters=512, is_training=False)
            t2 = time.perf_counter()
            if rep:                                                 # first pass = warm-up
                ts_enc.append(t1 - t0)
                ts_dec.append(t2 - t1)
    enc_s, dec_s = statistics.median(ts_enc), statistics.median(ts_dec)
    line = {"config": 1, "workload": "BTS-NYU DenseNet-161 inference, 1 synthetic %dx%d image, random-init weights, CPU" % (a.height, a.width),
            "kind": "restated-reference CPU (torch): torchvision densenet161 taps + UNMODIFIED bts_decoder.py over oracle/tf_shim; not TensorFlow",
            "cores": cores, "reps": a.reps, "encoder_s": round(enc_s, 4), "decoder_s": round(dec_s, 4),
            "images_per_s": round(1.0 / (enc_s + dec_s), 4), "decoder_images_per_s": round(1.0 / dec_s, 4),
            "depth_shape": list(depth.shape), "depth_mean": float(depth.mean()), "torch": torch.__version__}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
