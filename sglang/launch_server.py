"""`python -m sglang.launch_server --model-path llama-3-8b --enable-semi-pd [...]` — same entry
point and flag names as the reference (python/sglang/launch_server.py:10-16)."""
import sys

from semi_pd_b200.launch_server import main

if __name__ == "__main__":
    main(sys.argv[1:])
