"""Where one flat-EA acquisition optimisation (BASELINE.json config 4 shape: 10k-candidate pool, 3 generations) spends
its time: every stage timed with a device synchronisation behind it.  Usage: python profiles/tools_ea_breakdown.py [pop] [n_eval]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from bosot_b200 import sot
from bosot_b200.encoding import Grammar, random_stress_trees
from bosot_b200.flat_search import FlatCandidateGenerator, FlatEvolutionaryOptimizer, neighbours_device

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
E = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda:0")
gram = Grammar.get("CKSHighDimSearchSpace", 5)
w = np.array([.2, .7, .4]) / 1.3
ev = sot.as_device_trees(random_stress_trees(E, gram, seed=1), dev)
eng = sot.AcquisitionEngine(ev, torch.randn(E, dtype=torch.float64), gram, w, 1.3, 0.8, 1e-3, -0.5)
gen = FlatCandidateGenerator(gram)
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)
for rep in range(4):
    if rep == 1: T.clear()
    np.random.seed(rep)
    t0 = time.perf_counter(); pool = gen.get_initial_for_evolutionary_opt(pop); tick("initial population (host, native)", t0)
    t0 = time.perf_counter(); population = torch.from_numpy(pool).to(dev); tick("H2D population", t0)
    values = None
    best_pinned = torch.zeros(2, dtype=torch.float64).pin_memory()
    for step in range(3):
        if step > 0:
            k = int(pop / 5)
            t0 = time.perf_counter(); idx, _ = sot.topk(values, k, want_values=False); tick("topk", t0)
            t0 = time.perf_counter(); survivors = population[idx]; tick("gather survivors", t0)
            t0 = time.perf_counter(); counts = survivors[:, :1].cpu().numpy(); tick("D2H node counts", t0)
            t0 = time.perf_counter(); parent, index = gen.draw_offspring_indices(counts, 4); tick("draw offspring (host, native RNG replay)", t0)
            t0 = time.perf_counter(); par_d = torch.from_numpy(parent).to(dev); idx_d = torch.from_numpy(index); tick("H2D parents", t0)
            t0 = time.perf_counter(); offspring = neighbours_device(survivors[par_d].contiguous(), idx_d, gram); tick("neighbours kernel (+ status check)", t0)
            t0 = time.perf_counter(); population = torch.cat([offspring, survivors]); tick("concat", t0)
        t0 = time.perf_counter(); values = eng.acquire_device(population)[2].contiguous(); tick("acquire_device", t0)
        t0 = time.perf_counter(); sot.argmax_first(values, out=best_pinned); torch.cuda.current_stream().synchronize(); tick("argmax", t0)
        t0 = time.perf_counter(); _ = population[int(best_pinned[1])].cpu().numpy(); tick("D2H best record", t0)
tot = sum(T.values()) / 3
for k_, v in T.items():
    print("%-45s %.3f ms per EA run" % (k_, v / 3))
print("total %.3f ms per EA run (pop %d, E %d)" % (tot, pop, E))
ea = FlatEvolutionaryOptimizer(pop, 4, gen, dev)
for _ in range(2): ea.maximize(lambda t: eng.acquire_device(t)[2], 3)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): ea.maximize(lambda t: eng.acquire_device(t)[2], 3)
torch.cuda.synchronize(); print("FlatEvolutionaryOptimizer.maximize: %.3f ms" % (200 * (time.perf_counter() - t0)))
