"""GPU test (-m gpu) of the null-padded FheString operations at PARAM_MESSAGE_2_CARRY_2_KS_PBS with real keys: decrypted results
against Rust's str semantics (host/padded.h; SURVEY.md 8(f) N4, "full FheString surface")."""
import numpy as np
import pytest

from fhe_string_bounty_b200.host import Program
from helpers import engine_params

pytestmark = pytest.mark.gpu


def pad(s: bytes, cap: int) -> bytes:
    return s + b"\0" * (cap - len(s))


def test_padded_string_ops_gpu(orc, keys_2_2):
    import fhe_string_bounty_b200 as F
    from oracle import radix as R
    p, ck, sk = keys_2_2
    eng = F.Engine(engine_params(p))
    eng.upload_ksk(sk.ksk)
    eng.upload_bsk_std(sk.bsk)
    prm = engine_params(p)
    cap = 24
    ws = b" \t\n\r\x0b\x0c"
    dec = ck.decrypt_message_and_carry
    for s in (b"  Hello Zama \t\n", b"", b"no-space", b" " * cap):
        enc = R.encrypt_string(ck, pad(s, cap))
        assert R.decrypt_radix(ck, Program("pstring_len", (cap,), params=prm).run(eng, enc)) == len(s)
        assert dec(Program("pstring_is_empty", (cap,), params=prm).run(eng, enc)[0]) == int(not s)
        for op, want in (("trim_start", s.lstrip(ws)), ("trim_end", s.rstrip(ws)), ("trim", s.strip(ws))):
            P = Program("pstring_" + op, (cap,), params=prm)
            assert R.decrypt_string(ck, P.run(eng, enc)) == pad(want, cap), (op, s)
        for pat in (b"  He", b"no-", b"ace", b"\t\n"):
            out = Program("pstring_strip_prefix", (cap,), clear=pat, params=prm).run(eng, enc)
            want = s[len(pat):] if s.startswith(pat) else s
            assert dec(out[0]) == int(s.startswith(pat)) and R.decrypt_string(ck, out[1:]) == pad(want, cap), (s, pat)
            out = Program("pstring_strip_suffix", (cap,), clear=pat, params=prm).run(eng, enc)
            want = s[:-len(pat)] if s.endswith(pat) else s
            assert dec(out[0]) == int(s.endswith(pat)) and R.decrypt_string(ck, out[1:]) == pad(want, cap), (s, pat)
    print(f"trim, capacity {cap}: {P.n_pbs} PBS in {len(P.level_widths)} levels, {P.last_ms():.1f} ms on device")
    ca, cb = 12, 8
    for a, b in ((b"abcabcab", b"ab"), (b"encrypted", b"crypt"), (b"encrypted", b"ted"), (b"abc", b"abd"), (b"", b""), (b"same", b"same"), (b"xy", b"xyz")):
        enc = np.concatenate([R.encrypt_string(ck, pad(a, ca)), R.encrypt_string(ck, pad(b, cb))])
        want = {"eq": a == b, "ne": a != b, "lt": a < b, "le": a <= b, "gt": a > b, "ge": a >= b, "contains": b in a,
                "starts_with": a.startswith(b), "ends_with": a.endswith(b)}
        for op, w in want.items():
            assert dec(Program("pstring_" + op, (ca, cb), params=prm).run(eng, enc)[0]) == int(w), (op, a, b)
        assert R.decrypt_string(ck, Program("pstring_concat", (ca, cb), params=prm).run(eng, enc)) == pad(a + b, ca + cb)
        for op, pos in (("find", a.find(b)), ("rfind", a.rfind(b))):
            out = Program("pstring_" + op, (ca, cb), params=prm).run(eng, enc)
            assert (dec(out[0]), R.decrypt_radix(ck, out[1:])) == (int(pos >= 0), max(pos, 0)), (op, a, b)
    enc = R.encrypt_string(ck, pad(b"ab", 4))
    assert R.decrypt_string(ck, Program("pstring_repeat", (4, 3), params=prm).run(eng, enc)) == pad(b"ababab", 12)
    eng.close()
