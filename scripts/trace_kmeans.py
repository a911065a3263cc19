"""Timeline of the tensor-core k-means kernel (cycles): CTA 0's last tile of each pass, the
finalising CTA's sub-phases, and the spread over all CTAs.  Needs a trace build:
    make -C spml_b200/csrc clean; make -C spml_b200/csrc EXTRA=-DSPML_KM_TRACE"""
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
tr, fin, cta = (ctypes.c_longlong * 256)(), (ctypes.c_longlong * 256)(), (ctypes.c_longlong * (160 * 12 * 4))()
lib.spml_debug_kmtc_trace.argtypes = [ctypes.c_void_p] * 3
assert lib.spml_debug_kmtc_trace(tr, fin, cta) == 0
t = torch.tensor(list(tr)).view(16, 16)
names = {1: 'tile loaded', 2: 'flag seen', 3: 'tma+mma issued', 4: 't_full', 5: 'epilogue', 6: 'merged',
         7: 'rechecked', 11: 'ranked', 8: 'accumulated', 9: 'done counted', 10: 'published'}
print(name, '(cycles since the start of the pass; namb = rows re-scored exactly)')
for it in range(11):
  r = t[it]
  ev = ' '.join('%s@%d' % (names[k], int(r[k] - r[0])) for k in (1, 2, 3, 4, 5, 6, 7, 11, 8, 9, 10) if r[k] != 0)
  nxt = int(t[it + 1][0] - r[0]) if it < 10 else 0
  print('it %2d  %s  namb=%d  next pass starts @%d' % (it, ev, int(r[15]), nxt))
c = torch.tensor(list(cta)).view(160, 12, 4)
c = c[c[:, 1, 1] != 0]
print('%d CTAs traced (last tile of each pass); cycles' % c.shape[0])
for it in range(1, 10):
  wait = (c[:, it, 1] - c[:, it, 0]).float()
  work = (c[:, it, 2] - c[:, it, 1]).float()
  print('it %2d  flag wait min/mean/max %6d %6d %6d   flag->counted min/mean/max %6d %6d %6d' % (
      it, wait.min(), wait.mean(), wait.max(), work.min(), work.mean(), work.max()))
f = torch.tensor(list(fin)).view(16, 16)
print('finalising CTA (last image finalised in each pass): cycles after it counted itself last')
for it in range(10):
  r = f[it]
  print('it %2d  fence@%d  sums staged@%d  normalised+written@%d  fence@%d  published@%d' % (
      it, int(r[11] - r[10]), int(r[12] - r[10]), int(r[13] - r[10]), int(r[14] - r[10]), int(r[15] - r[10])))
