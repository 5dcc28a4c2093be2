def C(value, epoch, global_step, interpolation="linear"):
    """threestudio's scheduled-value helper: numbers pass through, [start_step, start, end, end_step]
    lists interpolate on the step (configs/gaussian_splatting.yaml:24); "exp" = log-linear."""
    if isinstance(value, (int, float)):
        return value
    value = list(value)
    if len(value) == 3:
        value = [0] + value
    start_step, start_value, end_value, end_step = value
    cur = global_step if isinstance(end_step, int) else epoch
    t = max(min(1.0, (cur - start_step) / max(end_step - start_step, 1e-9)), 0.0)
    if interpolation == "exp":
        import math
        return math.exp(math.log(start_value) * (1 - t) + math.log(end_value) * t)
    return start_value + (end_value - start_value) * t
