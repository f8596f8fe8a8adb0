import numpy as np, sys
tr=np.load(sys.argv[1])
tot=[]
for cta in range(0,148,2):
    pr=tr[cta,0]; pr=pr[pr>0]; p=pr[1:]; m=len(p)//2; pe=p[:m*2].reshape(m,2)
    tot.append((pe[:,1]-pe[:,0]).mean())
mma_tot=[]
for cta in range(0,148,2):
    mma=tr[cta,1]; mma=mma[mma>0]; n=len(mma)//4; ev=mma[:n*4].reshape(n,4)
    mma_tot.append(((ev[-1,3]-ev[0,0])/n, (ev[:,2]-ev[:,1]).mean(), (ev[:,1]-ev[:,0]).mean()))
mma_tot=np.array(mma_tot)
print(sys.argv[1], 'producer flag-check cycles per unit: mean %.0f' % np.mean(tot), '| issuer per unit: span %.0f operand-wait %.0f acc-wait %.0f' % tuple(mma_tot.mean(0)))
