import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.util import make_both, err_metrics
from pysdr_b200.bank import ReceiverBank
from pysdr_b200.receiver import receiver_offsets
from pysdr_b200.synth import synth_iq
from oracle import receiver_oracle as rxo
fcs=[-500,700,1400]; modes=['AM','NFM','USB']
P,Po=make_both(8,fcs,modes,af_bw_khz=[5,10,2])
offs=receiver_offsets(P)
n=3*P.IN_CHUNK_SIZE
xd=synth_iq(n,P.SRATE,offs,modes,seed=5,device='cuda')
x=xd.cpu().numpy()
rxo.create_receivers(Po)
ref=[np.concatenate([Po.rx[r].demod_data(x[c*P.IN_CHUNK_SIZE:(c+1)*P.IN_CHUNK_SIZE]) for c in range(3)]) for r in range(3)]
for direct in (False, True):
    b=ReceiverBank(P,offs,max_in=n)
    b.force_direct_fir(direct)
    am,iq,dc=b.process(xd)
    for r in range(3):
        g=am[r].cpu().numpy()
        print('direct' if direct else 'fft', r, modes[r], 'max', np.abs(g).max(), 'ref max', np.abs(ref[r]).max(), err_metrics(g,ref[r]), b.agc_get(r))
