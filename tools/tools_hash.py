"""SHA-256 of a state_dict (shared by tools/make_golden.py and the tests)."""
import hashlib


def sd_hash(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()
