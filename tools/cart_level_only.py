"""One CART fit with the level-batched split search (for kernel launch lists): python tools/cart_level_only.py [n k depth]"""
import sys

sys.argv = sys.argv[:4] + ["--level-only"]
exec(open("tools/cart_level_bench.py").read())
