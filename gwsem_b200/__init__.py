"""gwsem_b200: B200-native engine for gwsem's per-SNP association loop."""
from .model import (MultigroupModel, buildItem, buildOneFac, buildOneFacRes, buildOneItem, buildTwoFac,  # noqa: F401
                    coefNames, flatten, mxRename, omxSetParameters)
from .gwas import (GWAS, buildAnalysesPlan, numAvailableRecords, prepareComputePlan, processSnpData,  # noqa: F401
                   run_plan)
from .report import isSuspicious, loadResults, loadSuspicious, signif, signifGxE  # noqa: F401
from .model import setupExogenousCovariates, setupThresholds  # noqa: F401
