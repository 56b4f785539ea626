"""Asset lookup: the reference opens assets relative to the CWD (repo root); fall back to this package."""
import os

ROOT = os.path.dirname(os.path.abspath(__file__))


def resolve(path):
    path = path.replace("\\", "/")
    if os.path.exists(path):
        return path
    alt = os.path.join(ROOT, path)
    return alt if os.path.exists(alt) else path
