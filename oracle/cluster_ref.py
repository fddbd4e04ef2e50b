"""Literal pandas restatement of the selection rule inside the reference's `Adaptive_clustering`
(SpaDOT/utils/_analyze_utils.py:73-88).  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import pandas as pd


def select_n_clusters_ref(wss, min_clusters=4, wss_threshold=0.1):
    wss = list(wss)
    max_clusters = min_clusters + len(wss) - 1
    wss_diff = -np.diff(wss)                                                                   # :73
    wss_diff_ratios = [wss_diff[i] / wss_diff[i + 1] for i in range(len(wss_diff) - 1)]        # :74
    wss_df = pd.DataFrame({                                                                    # :75-80
        'clusters': range(min_clusters, max_clusters + 1),
        'wss': wss,
        'wss_diff': [None] + list(wss_diff),
        'wss_diff_ratio': [None] + list(wss_diff_ratios) + [None]
    })
    wss_range = wss_df['wss'].max() - wss_df['wss'].min()                                      # :81
    wss_diff_threshold = wss_threshold * wss_range                                             # :82
    filtered_wss_df = wss_df[wss_df['wss_diff'] > wss_diff_threshold]                          # :83
    max_idx = filtered_wss_df['wss_diff_ratio'].idxmax()                                       # :84
    return int(filtered_wss_df['clusters'][max_idx])                                           # :85
