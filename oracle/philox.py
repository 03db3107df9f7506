"""Counter-based Philox4x32-10 uniform generator (numpy) -- TEST INFRASTRUCTURE.

Part of the CPU oracle: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.  The CUDA kernels implement the identical generator
(csrc/philox.cuh) so that oracle and device draw the same Gumbel noise without shipping
noise tensors.

The reference draws its uniforms inside tensorflow_probability's
RelaxedOneHotCategorical.sample() (called at /root/reference/sga.py:97 and :120) from TF's own
Philox stream, which cannot be reproduced outside TF (op seeds derive from graph op ids,
sga.py:17).  We therefore define our own keyed stream:

    key     = (seed_lo, seed_hi)
    counter = (element_pair_index, step, tensor_id, 0)
    output  = 4 x uint32 -> uniforms for elements 2*pair and 2*pair+1, two logits each:
              (u0[e0], u1[e0], u0[e1], u1[e1])
    u = ((bits >> 8) + 0.5) * 2**-24   in (0, 1)   (never 0 or 1, so -log(-log u) is finite)
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds. Counters are uint32 arrays (broadcastable),
    keys python ints. Returns four uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64) + np.zeros_like(c0)
    c2 = np.asarray(c2, dtype=np.uint64) + np.zeros_like(c0)
    c3 = np.asarray(c3, dtype=np.uint64) + np.zeros_like(c0)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK32
        n1 = lo1
        n2 = (hi0 ^ c3 ^ np.uint64(k1)) & MASK32
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def bits_to_uniform(bits):
    """uint32 -> float32 uniform in (0,1): ((bits>>8)+0.5)*2^-24 (exact in float32)."""
    return ((bits >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


def sga_uniforms(numel, seed, step, tensor_id):
    """Return (u0, u1): two float32 arrays of length `numel` -- the uniforms behind the Gumbel
    noise of the DOWN and UP logits of every element (flat NHWC order)."""
    npair = (numel + 1) // 2
    idx = np.arange(npair, dtype=np.uint64)
    r0, r1, r2, r3 = philox4x32_10(idx & MASK32, np.uint64(step), np.uint64(tensor_id), np.uint64(0),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u0 = np.empty(npair * 2, dtype=np.float32)
    u1 = np.empty(npair * 2, dtype=np.float32)
    u0[0::2] = bits_to_uniform(r0)
    u1[0::2] = bits_to_uniform(r1)
    u0[1::2] = bits_to_uniform(r2)
    u1[1::2] = bits_to_uniform(r3)
    return u0[:numel], u1[:numel]


# Known-answer vectors of Philox4x32-10 from the Random123 distribution (kat_vectors):
#   ctr=0,0,0,0 key=0,0                      -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
#   ctr=ffffffff x4 key=ffffffff,ffffffff    -> 408f276d 41c83b0e a20bc7c6 6d5451fd
#   ctr=243f6a88 85a308d3 13198a2e 03707344 key=a4093822 299f31d0 -> d16cfe09 94fdcceb 5001e420 24126ea1
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
