import sys, time
sys.path.insert(0,'/root/repo/ff-lins_b200'); sys.path.insert(0,'/root/repo')
import numpy as np
import fflins
from fflins import api, synth
seq=synth.Sequence('robot')
c=api.Context(api.default_config(point_filter_num=10))
fr=seq.frame(0)
c.frame_process(1, fr['xyzi'], fr['time'], fr['ins_t'], fr['ins_p'], fr['ins_q'], seq.R_bl, seq.t_bl, fr['stamp'], fr['td'])
R,t=c.get_pose(1)
pose=api.Pose.make(R,t)
import ctypes as C
L=c.L; h=c.h
def bench(fn,n=2000):
    for _ in range(200): fn()
    t0=time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter()-t0)/n*1e6
ids=np.array([1],np.uint64); idp=ids.ctypes.data_as(C.c_void_p)
def launch_only(): L.fflins_reproject_window(h,1,idp,C.byref(pose))
def launch_sync(): L.fflins_reproject_window(h,1,idp,C.byref(pose)); L.fflins_synchronize(h)
def sync_only(): L.fflins_synchronize(h)
print('launch only us', bench(launch_only)); L.fflins_synchronize(h)
print('launch+sync us', bench(launch_sync))
print('sync only us', bench(sync_only))
def info(): 
    i=api.FrameInfo(); L.fflins_frame_info_get(h,C.c_uint64(1),C.byref(i))
print('trivial ctypes call us', bench(info))
