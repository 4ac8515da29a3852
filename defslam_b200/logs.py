"""On-disk result logs in the reference's formats, so that scripts/plotting.ipynb and Twiddle.py run
unchanged against this backend.

  Matches.txt          one line per tracked frame: "<%05u timestamp> <inliers> <outliers> <localPoints>"
                       (DefTracking.cc:321-328; opened at Thirdparty/ORBSLAM_2/src/Tracking.cc:150)
  ScaleVariation.txt   "<timestamp> <scale>" with C++ ostream default formatting of a double / float
                       (DefTracking.cc:507,568; Tracking.cc:151)
  ErrorGTs<%05u>.txt   per-point 3-D errors of one frame as Eigen prints a column MatrixXd:
                       6 significant digits, every coefficient right-aligned to the widest one
                       (GroundTruthFrame.cc:259-264 -> GroundTruthTools::saveResults, GroundTruthCalculator.cc:174-186)
"""
from __future__ import annotations

import os


def cxx_default(x: float) -> str:
    """std::ostream << double with default flags: %g with precision 6."""
    return "%g" % float(x)


def stamp5(timestamp: float) -> str:
    """std::internal << std::setfill('0') << std::setw(5) << (unsigned int) mTimeStamp"""
    return "%05u" % int(timestamp)


class ResultLogs:
    def __init__(self, output_path: str):
        self.path = output_path
        os.makedirs(output_path, exist_ok=True)
        self.matches = open(os.path.join(output_path, "Matches.txt"), "w")
        self.scalefile = open(os.path.join(output_path, "ScaleVariation.txt"), "w")

    def frame(self, timestamp: float, inliers: int, outliers: int, local_points: int) -> None:
        self.matches.write(f"{stamp5(timestamp)} {int(inliers)} {int(outliers)} {int(local_points)}\n")

    def scale(self, timestamp: float, scale: float) -> None:
        self.scalefile.write(f"{cxx_default(timestamp)} {cxx_default(scale)}\n")

    def errors(self, timestamp: float, errors) -> str:
        name = os.path.join(self.path, "ErrorGTs" + stamp5(timestamp) + ".txt")
        cells = [cxx_default(e) for e in errors]
        width = max((len(c) for c in cells), default=0)
        with open(name, "w") as f:
            f.write("\n".join(c.rjust(width) for c in cells))   # Eigen: rows joined by '\n', none after the last
        return name

    def close(self) -> None:
        self.matches.close()
        self.scalefile.close()
