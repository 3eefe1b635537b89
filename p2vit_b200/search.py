"""Coarse-to-fine mixed-precision search: the heaviest *caller* of the quantized forward.

Reference: test_quant.py:316-463.  Three stages, all host logic around `validate`:
  1. `sample_candidates`     random {4,8} configs under the 1.1 x 4-bit MAC budget (test_quant.py:323-341): first layer
                             8 bit, consecutive layer pairs share a width, the head draws its own;
  2. `rank_by_sensitivity`   omega = sum_i sensitivity_i * distance_i[bit_i]  (:343-373) with `global_distance` (per layer
                             weight-quantisation MSE per bit choice, collected at calibration) and a per-layer Hessian-trace
                             vector; candidates sorted by omega;
  3. `evolutionary_search`   population 25, 8 iterations of 10 mutations + 10 crossovers under the budget (:395-460).

`evaluate(bit_config) -> top-1 %` is the only thing that touches the GPU; `mixed_precision_search` builds it from
`runner.validate`, and the engine keeps one packed int8/int4 weight set + CUDA graph per distinct bit_config, so switching
between configurations costs a dictionary lookup (engine.py: VitEngine.plan).

Deliberate differences from the reference, selectable with `replicate_reference=True`:
  * the reference appends a child that violates the budget (or repeats) with the *previous* child's accuracy
    (`val_prec1` is simply not reassigned, :418-425); here such children are skipped;
  * the reference draws candidate widths uniformly and rejects almost all of them against the budget; here the draw
    probability is matched to the budget (sample_candidates: p_high), and `max_draws` bounds the loop;
  * global_distance columns: the int4 / int8 entries instead of the first two (distance_columns).
Every evaluation is memoised: the reference re-validates configurations it has already seen.
"""
import random


def model_cost(flops, bit_config):
    """sum_i MACs_i * bits_i  (test_quant.py:335)"""
    return sum(f * b for f, b in zip(flops, bit_config))


def budget(flops, ratio=1.1, base_bits=4):
    """test_quant.py:323"""
    return ratio * sum(f * base_bits for f in flops)


def sample_candidates(flops, rng, bit_choice=(4, 8), ratio=1.1, max_candidates=51, max_draws=1 << 20, p_high=None):
    """test_quant.py:324-341.  len(flops) = 1 + 2k + 1 layers (patch embed, k pairs, head): 50 for a 12-block ViT
    (qkv, proj | fc1, fc2 per block), 98 for ViT-L.
    `p_high`: probability of drawing the larger width for a pair.  The reference draws uniformly (0.5), which puts the expected
    cost at 1.5 x the 4-bit cost: about one draw in 3e5 fits the 1.1 x budget of a 12-block ViT and its loop spins for minutes
    to collect 51 configurations.  The default (None) matches the expected cost to the budget instead
    (p_high = (ratio - 1) * lo / (hi - lo), 0.1 for {4, 8}), same support, ~half of the draws accepted."""
    n = len(flops)
    lo, hi = min(bit_choice), max(bit_choice)
    limit = budget(flops, ratio, lo)
    if p_high is None:
        p_high = min(0.5, max(0.02, (ratio - 1.0) * lo / float(hi - lo)))
    assert len(bit_choice) == 2 or p_high == 0.5, "non-uniform draws are defined for two widths"
    draw = (lambda: rng.choice(bit_choice)) if p_high == 0.5 else (lambda: hi if rng.random() < p_high else lo)
    out = []
    for _ in range(max_draws):
        pairs = [draw() for _ in range(n // 2 - 1)]
        cfg = [hi] + [b for b in pairs for _ in range(2)] + [draw()]
        if not model_cost(flops, cfg) > limit and cfg not in out:
            out.append(cfg)
        if len(out) >= max_candidates:
            break
    return out


def rank_by_sensitivity(candidates, global_distance, sensitivity, bit_choice=(4, 8)):
    """test_quant.py:343-373: omega(cfg) = sum_{i>=1} sensitivity[i-1] * global_distance[i-1][index of cfg[i] in bit_choice].
    `global_distance[i]` holds one distance per bit choice for layer i+1 (the patch embed has none).  Returns
    [(cfg, omega)] sorted by omega, smallest first."""
    ranked = []
    for cfg in candidates:
        assert len(cfg) - 1 == len(global_distance) == len(sensitivity), (len(cfg), len(global_distance), len(sensitivity))
        omega = 0.0
        for i in range(1, len(cfg)):
            omega += float(sensitivity[i - 1]) * float(global_distance[i - 1][bit_choice.index(cfg[i])])
        ranked.append((cfg, omega))
    ranked.sort(key=lambda t: t[1])
    return ranked


def evolutionary_search(evaluate, seeds, flops, rng, bit_choice=(4, 8), ratio=1.1, pop_size=25, evo_iter=8, mutate_size=10,
                        mutate_prob=0.5, crossover_size=10, crossover_prob=0.5, replicate_reference=False, log=None):
    """test_quant.py:395-460.  `seeds`: configurations in ranked order (the first pop_size form the initial population).
    Returns the final population [(cfg, top1)] best first and the number of distinct configurations evaluated."""
    limit = budget(flops, ratio, min(bit_choice))
    cache = {}

    def score(cfg):
        key = tuple(cfg)
        if key not in cache:
            cache[key] = float(evaluate(list(cfg)))
        return cache[key]

    popu = [[list(c), score(c)] for c in seeds[:pop_size]]
    popu.sort(key=lambda t: t[1], reverse=True)
    for it in range(evo_iter):
        children, last = [], (popu[0][1] if popu else 0.0)
        seen = []
        while len(seen) <= mutate_size:                       # the reference's `> mutate_size` break: size + 1 children
            old = rng.choice(popu)[0]
            new = [b if rng.random() < mutate_prob else rng.choice(bit_choice) for b in old]
            ok = not model_cost(flops, new) > limit and new not in seen
            if ok:
                last = score(new)
            seen.append(new)
            if ok or replicate_reference:
                children.append([new, last])
        seen = []
        guard = 0
        while len(seen) <= crossover_size:
            a, b = rng.choice(popu)[0], rng.choice(popu)[0]
            guard += 1
            if a == b:
                if guard > 10000:                            # a population of identical parents cannot cross over
                    break
                continue
            new = [x if rng.random() < crossover_prob else y for x, y in zip(a, b)]
            ok = not model_cost(flops, new) > limit and new not in seen
            if ok:
                last = score(new)
            seen.append(new)
            if ok or replicate_reference:
                children.append([new, last])
        for child in children:
            if child[1] > popu[-1][1] and (replicate_reference or child[0] not in [p[0] for p in popu]):
                popu.append(child)
        popu.sort(key=lambda t: t[1], reverse=True)
        popu = popu[:pop_size]
        if log:
            log("evolution %d: best %.3f %%, worst kept %.3f %%, %d configurations evaluated" % (it, popu[0][1], popu[-1][1], len(cache)))
    return [(c, s) for c, s in popu], len(cache)


def distance_columns(global_distance, bit_choice=(4, 8), replicate_reference=False):
    """Each row of `global_distance` holds the weight-quantisation distance of one layer for every calibrated bit type in
    BIT_TYPE_LIST order without uint8 (uint3, uint4, int4, int8: layers.py:178-201).  The reference indexes the row with the
    position of the width in `bit_choice` (test_quant.py:351-354), i.e. reads the uint3 / uint4 entries for 4 / 8 bit (Q6);
    the fixed mapping takes the int4 / int8 entries."""
    width = len(global_distance[0])
    if replicate_reference or width == len(bit_choice):
        cols = list(range(len(bit_choice)))
    else:
        names = ["uint3", "uint4", "int4", "int8"][-width:] if width <= 4 else None
        assert names is not None, "unexpected global_distance row of %d entries" % width
        cols = [names.index("int%d" % b) for b in bit_choice]
    return [[float(row[c]) for c in cols] for row in global_distance]


def mixed_precision_search(model, flops, global_distance, val_batches, sensitivity=None, seed=0, bit_choice=(4, 8), top_validate=5,
                           ratio=1.1, replicate_reference=False, log=None, **evo):
    """The reference's `--mixed` flow on a calibrated + quantized model.  `flops`, `global_distance`: what the calibration
    forward returned (`runner.calibrate_model(model, images)[0][1:]`, as test_quant.py:306-309 keeps them).  Steps: candidates,
    sensitivity ranking, validation of the `top_validate` best by omega (test_quant.py:376-391), evolutionary search.
    `sensitivity`: per-layer Hessian traces (the reference hard-codes them, test_quant.py:207-259); None weighs every layer
    equally.  Returns a dict with the ranked list, the validated head of it and the final population."""
    from . import runner
    rng = random.Random(seed)
    flops = [float(f) for f in flops]
    gdist = distance_columns(global_distance, bit_choice, replicate_reference)
    if sensitivity is None:
        sensitivity = [1.0] * (len(flops) - 1)
    cands = sample_candidates(flops, rng, bit_choice, ratio, p_high=0.5 if replicate_reference else None, max_draws=1 << 16)
    ranked = rank_by_sensitivity(cands, gdist, sensitivity, bit_choice)

    def evaluate(cfg):
        return runner.validate(model, val_batches, cfg)[1]

    head = [(cfg, om, evaluate(cfg)) for cfg, om in ranked[:top_validate]]
    if log:
        for cfg, om, acc in head:
            log("omega %.4g  top-1 %.3f %%  %s" % (om, acc, "".join(str(b) for b in cfg)))
    popu, n_eval = evolutionary_search(evaluate, [c for c, _ in ranked], flops, rng, bit_choice, ratio, replicate_reference=replicate_reference,
                                       log=log, **evo)
    return {"flops": flops, "ranked": ranked, "validated": head, "population": popu, "evaluated": n_eval}
