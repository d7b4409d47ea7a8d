import ctypes as C, sys, os
sys.path.insert(0, "/root/repo")
from jaxfun_b200 import _lib as L
lib = L.load()
d = C.c_double()
L.check(lib.jfx_calibrate_dmma(None, 20000, C.byref(d)))
print(os.environ.get("JFX_CAL_BLOCKS_PER_SM"), "blocks/SM -> DMMA regs loop", d.value, "TFLOP/s")
