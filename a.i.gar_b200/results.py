"""The reference's result-file formats, fed from batched GPU episodes.

The reference evaluates a policy by running `n_tests` independent episodes of RESET_LIMIT frames, one process each
(src/aigar.py:523-538 performAgarioTest, :541-581 testAgarioModel), reduces Bot.totalMasses to mean / max mass per episode and
writes
  * `data/<name>MassOverTime.txt`  — one value per line, appended (exportResults :459-465, exportTestResults :499-519);
  * `final_results.txt`            — "Number of runs per testing: N" and one line per test type
                                     "<name> Highscore: .. Mean: .. StdMean: .. Mean_Max_Score: .. Std_Max_Score: .." (:609-627).
Here an "episode" is one env of an AgarBatch: the per-agent statistics the step kernels keep (sum / max / count of total mass,
bot.py:253 -> AGAR_GET_STATS) are exactly what those reductions need, so n_tests episodes cost one batched rollout.  File
contents are byte-for-byte what the reference's functions write for the same numbers (tests/test_results.py calls both).
"""
import os

import numpy as np


def episode_evals(mean_masses, max_masses, name, plot_name):
    """testAgarioModel's reduction (aigar.py:567-581) of per-episode mean and max masses."""
    mean_masses, max_masses = np.asarray(mean_masses, dtype=np.float64), np.asarray(max_masses, dtype=np.float64)
    return {"name": name, "plotName": plot_name, "meanScore": np.mean(mean_masses), "stdMean": np.std(mean_masses),
            "meanMaxScore": np.mean(max_masses), "stdMax": np.std(max_masses), "maxScore": np.max(max_masses)}


def evals_from_batch(batch, name="test", plot_name="Test", agent=0):
    """One evaluation entry from the episode statistics of every env of an AgarBatch (agent slot `agent`)."""
    from . import layout as lay
    stats = batch.get(lay.GET_STATS).cpu().numpy()[:, agent, :]  # [E][4]: sum, max, frames, deaths
    frames = np.maximum(stats[:, 2], 1.0)
    return episode_evals(stats[:, 0] / frames, stats[:, 1], name, plot_name)


def export_results(results, path, name):
    """exportResults (aigar.py:459-465): append one str(value) per line to <path><name>.txt."""
    with open(path + name + ".txt", "a") as f:
        for val in results:
            f.write(str(val) + "\n")


# test type -> file stem of exportTestResults (aigar.py:499-519)
_MASS_FILES = (("current", "testMassOverTime", lambda p: True),
               ("vsGreedy", "VS_1_GreedyMassOverTime", lambda p: p.get("MULTIPLE_BOTS_PRESENT")),
               ("virus", "Pellet_Collection_Virus_MassOverTime", lambda p: p.get("VIRUS_SPAWN")),
               ("virusGreedy", "VS_1_Greedy_Virus_MassOverTime", lambda p: p.get("VIRUS_SPAWN") and p.get("MULTIPLE_BOTS_PRESENT")))


def export_test_results(test_results, path, parameters):
    """exportTestResults without the matplotlib part: test_results is the list (one entry per test interval) of
    {test type: evals}; parameters a dict with MULTIPLE_BOTS_PRESENT / VIRUS_SPAWN."""
    data = os.path.join(path, "data") + os.sep
    if not os.path.exists(data):
        os.mkdir(data)
    for key, stem, enabled in _MASS_FILES:
        if enabled(parameters):
            export_results([val[key]["meanScore"] for val in test_results], data, stem)


def write_final_results(evals, path, n_runs):
    """runFinalTests (aigar.py:609-627): <path>/final_results.txt."""
    data = "Number of runs per testing: " + str(n_runs) + "\n"
    for test_type in evals:
        e = evals[test_type]
        data += (e["name"] + " Highscore: " + str(round(e["maxScore"], 1)) + " Mean: " + str(round(e["meanScore"], 1)) +
                 " StdMean: " + str(round(e["stdMean"], 1)) + " Mean_Max_Score: " + str(round(e["meanMaxScore"], 1)) +
                 " Std_Max_Score: " + str(round(e["stdMax"], 1)) + "\n")
    with open(os.path.join(path, "final_results.txt"), "w") as f:
        f.write(data)
    return data
