"""aurora_b200 — Blackwell (sm_100a) implementation of Aurora's forward pass behind the reference's API."""
