import sys, time, importlib, os, subprocess
code = '''
import sys, time, importlib, os
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests")
from synth import synth_local_ba
pkg=importlib.import_module("self_commit_orb-slam2_b200")
d=synth_local_ba()
nb=32
opt=pkg.Optimizer(max_kf=64,max_mp=5000,max_edges=30000,max_batch=nb)
opt.LocalBundleAdjustmentBatch([d]*nb)
t=time.perf_counter(); out=opt.LocalBundleAdjustmentBatch([d]*nb); dt=time.perf_counter()-t
print("chunk",os.environ.get("B2S_BA_CHUNK"),"ncta",os.environ.get("B2S_BA_NCTA"),"batch",nb,"ms %.2f"%(dt*1e3), flush=True)
'''
for chunk, ncta in [(32,4),(32,2),(16,8),(16,4),(8,16),(8,8),(11,13)]:
    env=dict(os.environ, B2S_BA_CHUNK=str(chunk), B2S_BA_NCTA=str(ncta), B2S_DEBUG_TIMING="1")
    subprocess.run([sys.executable,"-c",code],env=env)
