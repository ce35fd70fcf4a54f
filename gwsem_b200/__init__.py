"""gwsem_b200: B200-native engine for gwsem's per-SNP association loop."""
from .model import buildItem, buildOneFac, buildOneFacRes, buildOneItem, buildTwoFac, flatten  # noqa: F401
from .gwas import GWAS, buildAnalysesPlan, numAvailableRecords, processSnpData  # noqa: F401
from .report import isSuspicious, loadResults, signif, signifGxE  # noqa: F401
