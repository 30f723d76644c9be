"""Empty stand-in (oracle shim)."""
def parse_arguments(*a, **k): raise RuntimeError("isaacgym shim: no CLI")
def _unavailable(*a, **k): raise RuntimeError("isaacgym shim: simulator helpers are not available")
get_property_setter_map = get_property_getter_map = get_default_setter_args = _unavailable
apply_random_samples = check_buckets = generate_random_samples = _unavailable
