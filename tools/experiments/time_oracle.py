import time, torch, numpy as np, sys
from oracle import inside_oracle as O
N,L=int(sys.argv[1]),int(sys.argv[2])
W,T,pr=O.synthetic_grammar(N,32,0); tok,le=O.synthetic_tokens(1,L,32,1)
t=time.time(); lz,gW,gT,gP=O.flat_inside_with_grads(W,T,pr,tok[:,:,None],le); print(N,L,"inside+outside s:",time.time()-t, float(lz[0]), torch.get_num_threads())
