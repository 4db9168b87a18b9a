"""Host-to-device copy rate of the library's copy path: pinned, pageable through the driver, pageable through the pinned
ring (threads / chunk size from the environment).  python tools/h2d_probe.py [MiB ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import porla_b200 as pb
lib = pb.load(); lib.porla_device_init()
sizes = [int(x) for x in sys.argv[1:]] or [8, 24, 72]
tag = "threads=%s chunkKB=%s ring=%s" % (os.environ.get("PORLA_COPY_THREADS", "default"), os.environ.get("PORLA_COPY_CHUNK_KB", "1024"),
                                         "off" if os.environ.get("PORLA_NO_COPY_RING") else "on")
for mb in sizes:
    nbytes = mb << 20
    a = np.random.randint(0, 255, nbytes, dtype=np.uint8)
    pin = torch.from_numpy(a).pin_memory()
    r_pin = lib.porla_debug_h2d_rate(pin.data_ptr(), nbytes, 10)
    r_page = lib.porla_debug_h2d_rate(a.ctypes.data, nbytes, 10)
    print("%-40s %3d MiB: pinned %5.1f GB/s, pageable %5.1f GB/s" % (tag, mb, r_pin, r_page), flush=True)
