"""Golden for srfrd_b200.utils.df_data_partition: run the UNMODIFIED reference function (utils.py:92-139, exec'd
from the upstream checkout because utils.py as a whole does not parse, SURVEY.md note 4) on a small seeded frame and
store the frame plus its output.  TEST INFRASTRUCTURE; needs /root/reference (authoring container only)."""
import io, json, os, contextlib
from collections import defaultdict

import numpy as np
import pandas as pd

REF = os.environ.get("SRFRD_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "partition.json")

src = open(os.path.join(REF, "utils.py")).read().splitlines()
ns = {"defaultdict": defaultdict, "np": np}
exec("\n".join(src[91:139]), ns)                       # def df_data_partition ... return

rng = np.random.default_rng(7)
rows = []
for u in rng.permutation(np.arange(1, 41)):            # users appear in shuffled order, one of them only once
    n = 1 if u == 5 else int(rng.integers(2, 9))
    for t in range(n):
        rows.append(dict(user_id=int(u), item_id=int(rng.integers(1, 60)), time=t,
                         fake_review="fake" if rng.random() < 0.3 else "real"))
df = pd.DataFrame(rows)
res = {}
for valid in (False, True):
    with contextlib.redirect_stdout(io.StringIO()):
        tr, te, un, inum = ns["df_data_partition"](df, valid)
    res[str(valid)] = dict(train_items={str(k): [int(x) for x in v] for k, v in tr["item_ids"].items()},
                           train_reviews={str(k): [int(x) for x in v] for k, v in tr["review_ids"].items()},
                           test_items={str(k): [int(x) for x in v] for k, v in te["item_ids"].items()},
                           test_reviews={str(k): [int(x) for x in v] for k, v in te["review_ids"].items()},
                           usernum=int(un), itemnum=int(inum), order=[int(k) for k in tr["item_ids"].keys()])
json.dump(dict(frame=rows, expect=res), open(OUT, "w"))
print("wrote", OUT, len(rows), "rows")
