"""Stand-in for scikit-base (not installed here). Test infrastructure only: lets the
read-only reference at /root/reference be imported as the parity oracle."""
