"""python -m latteclip_b200.build  -- compile liblatte_b200.so for sm_100a (in-tree)."""
import sys

from ._lib import build

if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
