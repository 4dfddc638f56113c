"""Stand-in for the `timm` package (absent from this image, no network).

TEST INFRASTRUCTURE ONLY: lets /root/reference/layers/*win_attention.py import so the
live reference can pin the oracle.
"""
