import ctypes, torch, sys
sys.path.insert(0,'/root/repo')
from stoch_gpmp_b200 import _lib
lib=_lib.load(); dev=torch.device('cuda:0'); scratch=torch.zeros(16,device=dev)
st=ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
blocks=148*8
for mode,name,fl in ((5,'128 FFMA',128*2),(2,'64 FFMA2',64*4),(3,'128 FFMA + 64 LOP3',128*2),(4,'64 FFMA2 + 64 LOP3',64*4)):
    iters=2000
    lib.sgpmp_probe(mode,blocks,200,ctypes.c_void_p(scratch.data_ptr()),st)
    best=1e9
    for _ in range(3):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); lib.sgpmp_probe(mode,blocks,iters,ctypes.c_void_p(scratch.data_ptr()),st); e1.record(); torch.cuda.synchronize()
        best=min(best,e0.elapsed_time(e1))
    ninstr_fp = {5:128,2:64,3:192,4:128}[mode]
    print('%-30s %.3f ms  %.1f TFLOP/s  instr rate %.2f per clk per SMSP'%(name,best,blocks*256*iters*fl/(best*1e-3)/1e12, blocks*256*iters*ninstr_fp/32/(best*1e-3)/(148*4*1.965e9)))
