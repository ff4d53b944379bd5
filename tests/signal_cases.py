"""Golden vectors of the reference's own signal tests (tests/test_signal.nim) and doc examples
(impulse/signal.nim:545-594, 690-790) for the FFT-backed signal primitives.  Shared by the CPU tests
(host logic with a direct-convolution engine) and the GPU tests (product engine)."""
import numpy as np

T5 = np.array([1.0, 1.2, -0.3, -1.0, 2.0])

# tests/test_signal.nim:67-85
UPFIRDN = [
    (([1, 1, 1], [1, 1, 1]), {}, [1, 2, 3, 2, 1]),
    (([1, 2, 3], [1]), {"up": 3}, [1, 0, 0, 2, 0, 0, 3]),
    (([1, 2, 3], [1, 1, 1]), {"up": 3}, [1, 1, 1, 2, 2, 2, 3, 3, 3]),
    (([1.0, 1.0, 1.0], [0.5, 1.0, 0.5]), {"up": 2}, [0.5, 1.0, 1.0, 1.0, 1.0, 1.0, 0.5]),
    ((list(range(10)), [1]), {"down": 3}, [0, 3, 6, 9]),
    (([float(i) for i in range(10)], [0.5, 1.0, 0.5]), {"up": 2, "down": 3}, [0.0, 1.0, 2.5, 4, 5.5, 7.0, 8.5]),
]

# tests/test_signal.nim:87-147
RESAMPLE = [
    ({"up": 3}, [1.0006061736, 1.1652623795, 1.2269863677, 1.2007274083, 1.0003755977, 0.4963084977, -0.3001818521,
                 -1.1022817902, -1.4444776098, -1.0006061736, 0.1147783685, 1.3383072463, 2.0012123471, 1.7824018717,
                 0.9175789829]),
    ({"down": 2}, [0.9812293475, -0.0867375515, 0.5627008880]),
    ({"up": 3, "down": 2}, [1.0006061736, 1.2269863677, 1.0003755977, -0.3001818521, -1.4444776098, 0.1147783685,
                            2.0012123471, 0.9175789829]),
    ({"up": 6, "down": 4}, [1.0006061736, 1.2269863677, 1.0003755977, -0.3001818521, -1.4444776098, 0.1147783685,
                            2.0012123471, 0.9175789829]),
    ({"up": 3, "down": 2, "fir_order_factor": 30, "beta": 6.0},
     [1.0000890770, 1.1907035514, 1.0289575738, -0.3000267231, -1.4576402433, 0.1203579771, 2.0001781539, 0.9282917306]),
    ({"up": 3, "down": 2, "fir_order_factor": 0}, [1.0, 1.2, 1.2, -0.3, -1, -1, 2, 0]),
]
RESAMPLE_COMPLEX_IM = [2.0012123471, 2.4539727354, 2.0007511954, -0.6003637041, -2.8889552195, 0.2295567369, 4.0024246942,
                       1.8351579659]


def expected_resample_len(n, up, down):
    """ceil(n * up / down) with the rates reduced — what tests/test_signal.nim:150-250 tabulates."""
    import math
    g = math.gcd(up, down)
    up, down = up // g, down // g
    return n if up == down else math.ceil(n * up / down)
