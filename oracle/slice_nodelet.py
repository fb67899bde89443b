#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — build-time slicer for the reference's nodelet.

src/vofod_nodelet.cpp cannot be compiled whole here (ROS, mrs_lib, PCL, Eigen, OpenCV are absent), but the functions of
the per-scan path are plain C++ over a handful of types.  This script cuts those member functions and the small structs
they use out of the reference file WHERE IT LIES (/root/reference/src/vofod_nodelet.cpp) and writes them, verbatim, into
two generated include files under oracle/_ref/ (git-ignored — nothing of the reference is committed).
oracle/ref_nodelet_glue.cpp then includes them inside a stand-in `class VoFOD` whose members and helper types come from
oracle/shim/nodelet_shim.h, so the reference's own statements are what gets compiled and run.

usage: slice_nodelet.py <vofod_nodelet.cpp> <out_dir>
"""
import re
import sys

# (kind, signature regex).  The signature must match at the start of a (stripped) line of the reference file.
TYPES = [
    r"struct xyz_lut_t\b",
    r"enum class cluster_class_t\b",
    r"struct aabb_t\b",
    r"struct obb_t\b",
    r"struct cluster_t\b",
    r"struct detection_t\b",
    r"enum class profile_routines_t\b",
]
MEMBERS = [
    r"void initialize_apriori_map\(",                                    # :305-353  apriori map ingest (N3)
    r"void initialize_sensor_lut_simulation\(",                           # :374-420  sim XYZ LUT (A5)
    r"std::vector<uint8_t> load_mask\(",                                  # :506-560  sensor mask + mangle (A4 / N4)
    r"void processMsg\(const sensor_msgs::Range::ConstPtr msg\)",         # :580-613  rangefinder seed (A23)
    r"pc_XYZR_t::Ptr filterAndTransform\(",                               # :621-684  crop / transform / VoxelGridWeighted (A12, A10)
    r"std::vector<pcl::PointIndices> clusterCloud\(",                     # :689-698  (A13)
    r"std::pair<std::vector<pcl::PointIndices::ConstPtr>, std::vector<pcl::PointIndices::ConstPtr>> findCloseFarClusters\(",  # :703-750 (A14, A15)
    r"inline void updateVoxel\(",                                         # :777-797  (A11)
    r"void updateVMaps\(",                                                # :799-815  both overloads
    r"std::vector<cluster_t> classifyClusters\(",                         # :819-831
    r"std::vector<detection_t> extractDetections\(",                      # :834-879  (A19)
    r"void updateSeparatedBGClusters\(",                                  # :1126-1278 (A20, A21, A22)
    r"void raycast_cloud\(",                                              # :1397-1606 (A3, A6-A9)
    r"void reset\(\)",                                                    # :1610-1632 (A0)
    r"cluster_t classify_cluster\(",                                      # :1648-1731 (A16, A17)
]


def skip_ws_comments_strings(src, i):
    """index just past whatever non-code token starts at i (comment, string or char literal), or i if code starts there"""
    if src.startswith("//", i):
        j = src.find("\n", i)
        return len(src) if j < 0 else j
    if src.startswith("/*", i):
        j = src.find("*/", i + 2)
        return len(src) if j < 0 else j + 2
    if src[i] in "\"'":
        q = src[i]
        j = i + 1
        while j < len(src) and src[j] != q:
            j += 2 if src[j] == "\\" else 1
        return j + 1
    return i


def body_end(src, start):
    """index just past the '}' (and a directly following ';') that closes the first '{' at or after start"""
    i, depth, seen = start, 0, False
    while i < len(src):
        j = skip_ws_comments_strings(src, i)
        if j != i:
            i = j
            continue
        c = src[i]
        if c == "{":
            depth += 1
            seen = True
        elif c == "}":
            depth -= 1
            if seen and depth == 0:
                i += 1
                k = i
                while k < len(src) and src[k] in " \t":
                    k += 1
                if k < len(src) and src[k] == ";":
                    i = k + 1
                return i
        elif c == ";" and not seen:
            return -1  # a declaration, not a definition
        i += 1
    raise SystemExit("unbalanced braces after offset %d" % start)


def cut(src, line_starts, pattern):
    out = []
    rx = re.compile(r"^[ \t]*" + pattern, re.M)
    for m in rx.finditer(src):
        # code position?  (the reference has no such signature inside comments, but be safe: skip '//' and '*' lines)
        ls = src.rfind("\n", 0, m.start()) + 1
        head = src[ls:m.start() + 2].strip()
        if head.startswith("//") or head.startswith("*") or head.startswith("/*"):
            continue
        begin = ls
        # a template header on the previous line belongs to the definition
        prev_ls = src.rfind("\n", 0, ls - 1) + 1
        if src[prev_ls:ls].strip().startswith("template"):
            begin = prev_ls
        end = body_end(src, m.start())
        if end < 0:
            continue
        first_line = src.count("\n", 0, begin) + 1
        last_line = src.count("\n", 0, end) + 1
        out.append((first_line, last_line, src[begin:end]))
    if not out:
        raise SystemExit("slice_nodelet: no definition matches /%s/ — the reference file is not the one this script was written for" % pattern)
    return out


def main():
    ref, out_dir = sys.argv[1], sys.argv[2]
    src = open(ref).read()
    for name, pats in (("nodelet_types.inc", TYPES), ("nodelet_members.inc", MEMBERS)):
        parts = []
        for p in pats:
            for (a, b, text) in cut(src, None, p):
                parts.append("// ---- vofod_nodelet.cpp:%d-%d ----\n#line %d \"%s\"\n%s\n" % (a, b, a, ref, text))
        with open("%s/%s" % (out_dir, name), "w") as f:
            f.write("// GENERATED by oracle/slice_nodelet.py from %s — do not commit\n" % ref)
            f.write("\n".join(parts))
    return 0


if __name__ == "__main__":
    sys.exit(main())
