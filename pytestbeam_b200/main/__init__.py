"""Drop-in counterparts of the reference's ``pytestbeam/main`` package for the hot path (hit, device)."""
