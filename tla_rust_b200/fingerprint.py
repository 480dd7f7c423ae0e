"""Host-side mirror of tlag_fingerprint (csrc/tlag_vm.h) -- used only to route the (few) initial
states to their owner rank in multi-GPU runs."""
M64 = (1 << 64) - 1


def _rotl(x, r):
    return ((x << r) | (x >> (64 - r))) & M64


def _fmix(k):
    k ^= k >> 33
    k = (k * 0xff51afd7ed558ccd) & M64
    k ^= k >> 33
    k = (k * 0xc4ceb9fe1a85ec53) & M64
    k ^= k >> 33
    return k


def fingerprint_words(words) -> int:
    w = [int(x) & 0xFFFFFFFF for x in words]
    W = len(w)
    h = 0x9E3779B97F4A7C15 ^ ((W * 0xD6E8FEB86659FD93) & M64)
    i = 0
    while i + 1 < W:
        k = w[i] | (w[i + 1] << 32)
        k = (k * 0x87c37b91114253d5) & M64
        k = _rotl(k, 31)
        k = (k * 0x4cf5ad432745937f) & M64
        h ^= k
        h = (_rotl(h, 27) * 5 + 0x52dce729) & M64
        i += 2
    if i < W:
        k = w[i]
        k = (k * 0x87c37b91114253d5) & M64
        k = _rotl(k, 31)
        k = (k * 0x4cf5ad432745937f) & M64
        h ^= k
    h = _fmix(h ^ W)
    return h if h else 1


def owner_rank(fp: int, world: int) -> int:
    return (fp * world) >> 64


def owner_of_words(words, world: int, k: int = 2) -> int:
    """Mirror of tlag_owner_k (csrc/tlag_vm.h): owner = hash of the last k packed words, scaled to `world`."""
    w = [int(x) & 0xFFFFFFFF for x in words]
    W = len(w)
    key = (w[-1] << 32) | (w[-2] if W >= 2 else 0)
    h = _fmix((key * 0x9E3779B97F4A7C15 + 0x7F4A7C15) & M64)
    i, n = W - 3, 2
    while i >= 0 and n < k:
        h = _fmix(h ^ ((w[i] * 0x9E3779B97F4A7C15 + n) & M64))
        i -= 1
        n += 1
    return ((h >> 32) * world) >> 32
