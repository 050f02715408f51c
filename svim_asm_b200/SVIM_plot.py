"""sv-lengths.png (reference src/svim_asm/SVIM_plot.py) is cosmetic and needs matplotlib, which this image does
not ship: the CLI skips the plot with a log line when matplotlib is missing."""
import logging


def plot_sv_lengths(deletion_candidates, inversion_candidates, int_duplication_candidates, tan_dup_candidates,
                    novel_insertion_candidates, options):
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        logging.info("matplotlib is not installed: skipping sv-lengths.png")
        return
    lengths = {
        "DEL": [v.get_source()[2] - v.get_source()[1] for v in deletion_candidates],
        "INS": [v.get_destination()[2] - v.get_destination()[1] for v in novel_insertion_candidates],
        "INV": [v.get_source()[2] - v.get_source()[1] for v in inversion_candidates],
        "DUP_INT": [v.get_destination()[2] - v.get_destination()[1] for v in int_duplication_candidates],
        "DUP_TAN": [v.get_destination()[2] - v.get_destination()[1] for v in tan_dup_candidates],
    }
    fig, axes = plt.subplots(2, 1, figsize=(8, 6))
    for ax, (limit, step, log) in zip(axes, ((2000, 10, False), (20000, 100, True))):
        ax.hist([lengths[k] for k in lengths], bins=range(0, limit + step, step), stacked=True, label=list(lengths), log=log)
        ax.set_xlabel("SV length (bp)")
        ax.set_ylabel("count")
    axes[0].legend()
    fig.tight_layout()
    fig.savefig(options.working_dir + "/sv-lengths.png")
    plt.close(fig)
