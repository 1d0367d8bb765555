"""Entry-point shim: keeps the reference's module path `python -m sglang.launch_server` alive and
forwards to semi_pd_b200 (python/sglang/launch_server.py:1-16 in infinigence/Semi-PD)."""
