"""End-to-end compute_multi_exp with the caller's buffers pinned / pageable, over 1..k devices inside the call.

    python tools/e2e_pageable.py [log2n ...]          (PORLA_NO_COPY_RING=1: pageable copies through the driver's staging)

Pageable = ordinary heap memory (numpy arrays), what the reference's callers pass (`new[]` arrays, Client.hpp:124-127).
"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import porla_b200 as pb


def main():
    lib = pb.load()
    lib.porla_device_init()
    ndev_all = lib.porla_device_count()
    sizes = [int(x) for x in sys.argv[1:]] or [20]
    res = {"copy_ring": "off" if os.environ.get("PORLA_NO_COPY_RING") else "on", "devices_visible": ndev_all}
    for lg in sizes:
        n = 1 << lg
        g = torch.Generator(device="cuda")
        g.manual_seed(lg)
        ks = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
        tab = pb.Table.multiples_of_generator(pb.CURVE_BN254, ks.data_ptr(), n, pb.SCALAR_LE32, on_device=True)
        pts_np = np.frombuffer(tab.export(), dtype=np.uint8).copy()
        tab.destroy()
        sc_np = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)).numpy().copy()
        pts_pin = torch.from_numpy(pts_np).pin_memory()
        sc_pin = torch.from_numpy(sc_np).pin_memory()
        out = (C.c_ubyte * 64)()

        def call(sc_ptr, pt_ptr, ndev):
            lib.porla_msm_host_devices(pb.CURVE_BN254, C.c_void_p(sc_ptr), C.c_void_p(pt_ptr), n, pb.SCALAR_BE32, pb.POINT_BE64, ndev,
                                       C.cast(out, C.c_void_p))
            return bytes(out)

        def timeit(sc_ptr, pt_ptr, ndev, reps=8):
            lib.porla_measure_pint(1, 0.1)
            for _ in range(2):
                r = call(sc_ptr, pt_ptr, ndev)
            t = time.perf_counter()
            for _ in range(reps):
                call(sc_ptr, pt_ptr, ndev)
            return (time.perf_counter() - t) / reps * 1e3, r

        row = {}
        ref = None
        nd = 1
        while nd <= ndev_all:
            ms_pin, r1 = timeit(sc_pin.data_ptr(), pts_pin.data_ptr(), nd)
            ms_page, r2 = timeit(sc_np.ctypes.data, pts_np.ctypes.data, nd)
            ref = ref or r1
            assert r1 == ref and r2 == ref, "results differ"
            row["ndev%d" % nd] = {"pinned_ms": round(ms_pin, 3), "pageable_ms": round(ms_page, 3)}
            nd *= 2
        res["2^%d" % lg] = row
    print(json.dumps(res))


if __name__ == "__main__":
    main()
