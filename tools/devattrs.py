import torch, ctypes
cudart = ctypes.CDLL("libcudart.so.12")
def attr(a):
    v = ctypes.c_int(0); cudart.cudaDeviceGetAttribute(ctypes.byref(v), a, 0); return v.value
torch.cuda.init(); torch.zeros(1, device="cuda")
print("l2CacheSize", attr(38), "maxPersistingL2", attr(108), "maxAccessPolicyWindow", attr(109), "SMs", attr(16), "smemPerBlockOptin", attr(97), "coop", attr(95), "clusterLaunch", attr(120))
