"""Who waits in a warp-specialised kernel?  From an `ncu --set full --import-source on` report, print for every profiled launch the execution
counts of the single-thread roles' instructions: the mbarrier `try_wait`s (first try / retry loop) next to the TMA loads (UTMALDG) and MMAs
(UTCHMMA) they guard.  A retry count of the order of the first-try count means that role is the one that waits.
   python tools/ncu_roles.py <report.ncu-rep> [max_launches]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    i, n = 0, 0
    while i < len(rows) and n < limit:
        if rows[i] and rows[i][0] == "Kernel Name":
            print("##", rows[i][1])
            hdr = rows[i + 1]
            isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
            i += 2
            while i < len(rows) and not (rows[i] and rows[i][0] == "Kernel Name"):
                r = rows[i]
                if len(r) > iex and any(k in r[isrc] for k in ("TRYWAIT", "UTMALDG", "UTCHMMA", "UTCBAR", "UTMASTG")) and r[iex] not in ("0", ""):
                    print(f"   {r[isrc].strip()[:90]:92s} executed {r[iex]}")
                i += 1
            n += 1
        else:
            i += 1


if __name__ == "__main__":
    main()
