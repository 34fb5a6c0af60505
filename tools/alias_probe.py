import torch, sys
sys.path.insert(0, "/root/repo")
class M: pass
a = torch.zeros(16, device="cuda")
m = M(); m.__cuda_array_interface__ = {"shape": (16,), "typestr": "<f4", "data": (a.data_ptr(), False), "version": 2}
v = torch.as_tensor(m, device=torch.device("cuda", 0))
a += 1
torch.cuda.synchronize()
print("alias" if float(v[0]) == 1.0 else "COPY", v.data_ptr() == a.data_ptr())
