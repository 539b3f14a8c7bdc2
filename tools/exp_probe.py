import ctypes, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import basisu_rs_b200 as b
L = b.lib(); assert L.b2bu_init(0) == 0
a, m = ctypes.c_double(), ctypes.c_double()
print(L.b2bu_probe_int_peak(ctypes.byref(a), ctypes.byref(m)), "alu-pipe Tops/s", a.value, "alu+fma mix Tops/s", m.value)
