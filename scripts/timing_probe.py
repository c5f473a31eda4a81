import os,sys
sys.path.insert(0,'/root/repo/gimp-fix-ca_b200'); sys.path.insert(0,os.path.join(os.environ.get('GRAFT_REPO_ROOT','/root/repo'),'gimp-fix-ca_b200'))
import torch,fixca
KW = dict(blue=3.0, red=-2.0, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)
nf,h,w=32,2160,3840; bpp=3; pitch=(w*bpp+127)//128*128
src=torch.randint(0,255,(nf,h,pitch),dtype=torch.uint8,device='cuda'); dst=torch.empty_like(src)
p=fixca.FixCaParams(interpolation=2,lens_x=w//2,lens_y=h//2,**KW)
st=torch.cuda.current_stream().cuda_stream
fixca.fix_ca_frames_dev(src.data_ptr(),pitch,pitch*h,dst.data_ptr(),pitch,pitch*h,nf,w,h,bpp,1,p,fixca.PRECISION_FAST,st)
torch.cuda.synchronize()
print(fixca.last_kernel())
