"""Writes tests/golden/fet_kats.json.

The values are transcribed from the reference's own tests and README (file:line given per entry);
nothing is computed here -- the Haskell reference cannot run in this image (no ghc/stack).
"""
import json
from pathlib import Path

golden = {
    # /root/reference/test/unit/FETSpec.hs:13-100  (goa, gob, gta, gtb, mode, expected)
    "fetspec": [
        [2, 7, 5, 3, "two", 0.153434800493624, "FETSpec.hs:13-23"],
        [2, 7, 5, 3, "one", 0.11703002879473468, "FETSpec.hs:24-34"],
        [129, 173, 152, 138, "two", 2.1126884355092402e-2, "FETSpec.hs:35-45"],
        [129, 173, 152, 138, "one", 1.1267280239712145e-2, "FETSpec.hs:46-56"],
        [222, 211, 444, 555, "two", 1.7384585693541843e-2, "FETSpec.hs:57-67"],
        [222, 211, 444, 555, "one", 8.692292846770922e-3, "FETSpec.hs:68-78"],
        [99222, 98211, 77444, 74555, "two", 4.7092508444146475e-5, "FETSpec.hs:79-89"],
        [100998, 100555, 501, 555, "one", 4.193675848378964e-2, "FETSpec.hs:90-100"],
    ],
    # /root/reference/test/unit/ComparisonSpec.hs:43-51  (vector, char, indices, expected)
    "countspec": [
        ["00abcdefg", "0", [0, 1], 2, "ComparisonSpec.hs:44-47"],
        ["ty39193883939", "9", [3, 4, 5], 2, "ComparisonSpec.hs:48-51"],
    ],
    # /root/reference/README.md end-to-end rows: (file, mode, group1 desc, group2 desc, ratio_filter,
    #   rows in printed order [name, goa, gob, gta, gtb, "p", "ratio"])
    "readme": [
        ["test_binary.txt", "binary", "group B", "group C", 0.0,
         [["binary21", 4, 0, 0, 2, "1.0", "1.0"], ["binary44", 3, 1, 0, 2, "1.0", "0.75"],
          ["binary42", 3, 1, 0, 2, "1.0", "0.75"], ["binary24", 1, 3, 1, 0, "1.0", "-0.75"]],
         "README.md:116-125"],
        ["test_binary.txt", "binary", "position sideways", "position up", 0.0,
         [["binary9", 2, 0, 1, 4, "1.0", "0.8"], ["binary47", 2, 0, 1, 4, "1.0", "0.8"],
          ["binary8", 0, 2, 4, 1, "1.0", "-0.8"], ["binary49", 2, 0, 1, 4, "1.0", "0.8"],
          ["binary1", 0, 2, 3, 1, "1.0", "-0.75"]],
         "README.md:127-139"],
        ["test_binary.txt", "binary", "group B", "group C", 1.0,
         [["binary21", 4, 0, 0, 2, "1.0", "1.0"]],
         "README.md:170-180"],
        ["test_snps.txt", "snp", "group B", "group C", 0.0,
         [["snp32_a", 0, 4, 2, 0, "1.0", "-1.0"], ["snp45_a", 0, 4, 2, 0, "1.0", "-1.0"],
          ["snp41_t", 0, 4, 2, 0, "1.0", "-1.0"], ["snp21_t", 0, 4, 2, 0, "1.0", "-1.0"]],
         "README.md:214-222"],
    ],
    # values of an independent restatement (derived, not printed by feht)
    "derived_choose": [
        [60, 30, 1.1826458156486142e+17], [120, 49, 1.2930805958841516e+34],
        [120, 50, 1.8361744461555046e+34], [100, 50, 1.0089134454556308e+29],
        [281, 129, 7.224946190275846e+82], [311, 173, 2.6373259406807332e+91],
        [592, 302, 4.705645030527099e+176], [999, 499, 1.3514412047271443e+299],
    ],
    "derived_two_tail": [
        [1, 1, 1, 1, 0.9999999999999999], [3, 0, 0, 3, 0.1], [0, 5, 0, 7, 1.0], [5, 0, 7, 0, 1.0],
        [46, 0, 229, 5, 0.5952343051609723], [200, 200, 180, 220, 0.17852875822321457],
        [250, 249, 260, 240, 0.5691584133790883], [10, 489, 25, 475, 0.014847705943346461],
        [0, 499, 0, 500, 1.0], [395, 1, 3, 393, 1.1131536985421843e-227],
        [400, 0, 0, 400, 1.063589676951889e-239],
    ],
    "derived_chi2_p": [
        [2.0, 0.1572992070325112], [2.0000001, 0.1572991967209787],
        [3.841458820694124, 0.050000000015377855], [10.0, 0.0015654022584840055],
        [30.0, 4.3204630539861455e-08], [50.0, 1.5374368445009168e-12],
        [60.0, 9.43689570931383e-15], [70.0, 1.1102230246251565e-16], [73.0, 0.0],
    ],
    "derived_misc": {"logGamma_0.5": 0.5723649426171691, "logGamma_1.5": -0.12078223765633668,
                     "logBeta_291_303": -413.18893362428275},
    # raw (uncorrected) rows of config 1 from the same independent restatement
    "derived_c1_raw": [
        ["group A", "group B", "binary4", 0, 5, 3, 1, 0.047619047619047616, -0.75],
        ["group A", "group B", "binary21", 2, 4, 4, 0, 0.07619047619047618, -0.6666666666666667],
        ["group B", "group C", "binary21", 4, 0, 0, 2, 0.06666666666666667, 1.0],
        ["position sideways", "position up", "binary9", 2, 0, 1, 4, 0.14285714285714285, 0.8],
        ["position sideways", "position up", "binary8", 0, 2, 4, 1, 0.14285714285714285, -0.8],
    ],
}

Path(__file__).with_name("fet_kats.json").write_text(json.dumps(golden, indent=1) + "\n")
print("wrote fet_kats.json")
