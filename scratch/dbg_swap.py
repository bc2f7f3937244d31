import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.util import make_both, err_metrics
from pysdr_b200.bank import ReceiverBank
from pysdr_b200.receiver import receiver_offsets
from oracle import receiver_oracle as rxo
def noise(n, seed, scale=0.1):
    rng = np.random.default_rng(seed)
    return ((rng.normal(size=n) + 1j * rng.normal(size=n)) * scale).astype(np.complex64)
for variant in ('swap', 'retune', 'both', 'swap_generic', 'swap_1rx'):
    fcs = [1000, 1250] if variant != 'swap_1rx' else [1000]
    modes = ['USB', 'CW'][:len(fcs)]
    P, Po = make_both(8, fcs, modes, af_bw_khz=[2, .5][:len(fcs)])
    C = P.IN_CHUNK_SIZE
    x = noise(4 * C, 23)
    bank = ReceiverBank(P, receiver_offsets(P), max_in=2 * C)
    if variant == 'swap_generic': bank.force_generic(True)
    rxo.create_receivers(Po)
    am, iq, _ = bank.process(torch.from_numpy(x[:2 * C]).cuda())
    got = [[a.cpu().numpy().copy()] for a in am]; gi = [[a.cpu().numpy().copy()] for a in iq]
    if variant in ('retune', 'both'):
        bank.set_freq(len(fcs) - 1, 271828.1828); Po.rx[len(fcs) - 1].lo.change_freq(271828.1828)
    if variant != 'retune':
        bank.set_dec_taps(0, bank.filter_bank[4]); Po.rx[0].dec.h = Po.rx[0].dec.filter_bank[4]
    am, iq, _ = bank.process(torch.from_numpy(x[2 * C:]).cuda())
    for r in range(len(fcs)):
        got[r].append(am[r].cpu().numpy().copy()); gi[r].append(iq[r].cpu().numpy().copy())
        ref, refi = [], []
        for c in range(4):
            ref.append(Po.rx[r].demod_data(x[c * C:(c + 1) * C])); refi.append(Po.rx[r].iq.copy())
        print(variant, r, 'am', err_metrics(np.concatenate(got[r]), np.concatenate(ref)), 'iq', err_metrics(np.concatenate(gi[r]), np.concatenate(refi)))
