"""Per-iteration timeline of CTA 0 of the persistent k-means kernel
(build with `make -C spml_b200/csrc EXTRA=-DSPML_KM_TRACE`)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spml_b200 import _lib, synth
from spml_b200.static_head import StaticContrastiveHead
name = sys.argv[1] if len(sys.argv) > 1 else 'voc_scribble_b1'
w = synth.WORKLOADS[name]
head = StaticContrastiveHead(synth.make_config(w), w.batch, w.height, w.width, use_graph=False)
for s in range(3):
  b = {k: v.cuda() for k, v in synth.make_batch(w, step=s).items()}
  head.step(b['embedding'], b['semantic_label'], b['instance_label'], b['semantic_tag'], b['local_feature'])
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * 128)()
lib.spml_debug_km_trace.argtypes = [ctypes.c_void_p]
assert lib.spml_debug_km_trace(buf) == 0
t = torch.tensor(list(buf)).view(16, 8)
print(name)
for it in range(11):
  r = t[it]
  print('it %2d  pre-assign %6d  assign %6d (sync %d stage %d sync %d dots %d) accumulate %6d  barrier %6d  total %6d' % (
      it, int(r[1] - r[0]) if r[1] else 0, int(r[2] - r[1]) if r[1] else 0,
      int(r[5] - r[1]) if r[5] else 0, int(r[6] - r[5]) if r[5] else 0, int(r[7] - r[6]) if r[5] else 0, int(r[2] - r[7]) if r[5] else 0,
      int(r[3] - (r[2] if r[2] else r[0])), int(r[4] - r[3]), int(r[4] - r[0])))
