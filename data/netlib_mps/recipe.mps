NAME          RECIPE
ROWS
 E  BAL...BE
 E  BCC...BE
 E  BCH...BE
 E  BCL...BE
 E  BHH...BE
 E  BHL...BE
 E  BHX...BE
 E  BLV...BE
 E  BN4...BE
 E  BP8...BE
 E  B3E...BE
 E  B3P...BE
 E  B3R...BE
 E  B3T...BE
 E  B3E.VOBE
 E  B3P.VOBE
 E  B3R.VOBE
 G  NOM.3EBE
 G  NRO.3EBE
 G  N13.3EBE
 G  N14.3EBE
 G  N24.3EBE
 G  N36.3EBE
 G  NOM.3PBE
 G  NRO.3PBE
 G  N13.3PBE
 G  N14.3PBE
 G  N24.3PBE
 G  N36.3PBE
 G  NOM.3RBE
 G  NRO.3RBE
 G  N13.3RBE
 G  N14.3RBE
 G  N24.3RBE
 G  N36.3RBE
 E  WMO.3EBE
 E  WRO.3EBE
 E  WMO.3PBE
 E  WRO.3PBE
 E  WMO.3RBE
 E  WRO.3RBE
 L  XRV.3EBE
 L  X15.3EBE
 L  XRV.3PBE
 L  X15.3PBE
 L  XRV.3RBE
 L  X15.3RBE
 N  FAT...J.
 E  B&,1..BE
 E  BAL1..BE
 E  BCC1..BE
 E  BCH1..BE
 E  BCL1..BE
 E  BHH1..BE
 E  BHL1..BE
 E  BHX1..BE
 E  BLV1..BE
 E  BN41..BE
 E  BP81..BE
 E  B&,2..BE
 E  BAL2..BE
 E  BCC2..BE
 E  BCH2..BE
 E  BCL2..BE
 E  BHH2..BE
 E  BHL2..BE
 E  BHX2..BE
 E  BLV2..BE
 E  BN42..BE
 E  BP82..BE
 E  B&,3..BE
 E  BAL3..BE
 E  BCC3..BE
 E  BCH3..BE
 E  BCL3..BE
 E  BHH3..BE
 E  BHL3..BE
 E  BHX3..BE
 E  BLV3..BE
 E  BN43..BE
 E  BP83..BE
 E  B&,4..BE
 E  BAL4..BE
 E  BCC4..BE
 E  BCH4..BE
 E  BCL4..BE
 E  BHH4..BE
 E  BHL4..BE
 E  BHX4..BE
 E  BLV4..BE
 E  BN44..BE
 E  BP84..BE
COLUMNS
    BAL.3EBE  BAL...BE           -1.   B3E.VOBE            1.
    BAL.3EBE  XRV.3EBE           8.2   N13.3EBE            9.
    BAL.3EBE  N14.3EBE           12.   N24.3EBE           80.
    BAL.3EBE  N36.3EBE           99.   X15.3EBE           16.
    BAL.3EBE  WRO.3EBE      93.61705   WMO.3EBE       88.6782
    BCC.3EBE  BCC...BE           -1.   B3E.VOBE            1.
    BCC.3EBE  XRV.3EBE          21.1   N13.3EBE         116.8
    BCC.3EBE  N14.3EBE          114.   N24.3EBE          100.
    BCC.3EBE  N36.3EBE          100.   X15.3EBE          100.
    BCC.3EBE  WRO.3EBE      99.22401   WMO.3EBE      80.06283
    BCH.3EBE  BCH...BE           -1.   B3E.VOBE            1.
    BCH.3EBE  XRV.3EBE            2.   N13.3EBE         -14.8
    BCH.3EBE  N14.3EBE          -12.   N24.3EBE           2.4
    BCH.3EBE  N36.3EBE           90.   X15.3EBE          -8.2
    BCH.3EBE  WRO.3EBE      83.80122   WMO.3EBE      74.69736
    BCL.3EBE  BCL...BE           -1.   B3E.VOBE            1.
    BCL.3EBE  XRV.3EBE            8.   N13.3EBE            2.
    BCL.3EBE  N14.3EBE            9.   N24.3EBE           95.
    BCL.3EBE  N36.3EBE          100.   X15.3EBE           43.
    BCL.3EBE  WRO.3EBE      90.17511   WMO.3EBE      79.19421
    BHH.3EBE  BHH...BE           -1.   B3E.VOBE            1.
    BHH.3EBE  XRV.3EBE            1.   N13.3EBE         -21.9
    BHH.3EBE  N14.3EBE          -18.   N24.3EBE            4.
    BHH.3EBE  N36.3EBE          96.5   X15.3EBE         -12.5
    BHH.3EBE  WRO.3EBE       85.9962   WMO.3EBE      78.56822
    BHL.3EBE  BHL...BE           -1.   B3E.VOBE            1.
    BHL.3EBE  XRV.3EBE          12.5   N13.3EBE           37.
    BHL.3EBE  N14.3EBE           49.   N24.3EBE           98.
    BHL.3EBE  N36.3EBE          100.   X15.3EBE           65.
    BHL.3EBE  WRO.3EBE      86.96338   WMO.3EBE      82.92224
    BHX.3EBE  BHX...BE           -1.   B3E.VOBE            1.
    BHX.3EBE  XRV.3EBE            1.   N13.3EBE         -21.9
    BHX.3EBE  N14.3EBE          -18.   N24.3EBE            4.
    BHX.3EBE  N36.3EBE          96.5   X15.3EBE          -12.
    BHX.3EBE  WRO.3EBE       93.1476   WMO.3EBE      82.59274
    BLV.3EBE  BLV...BE           -1.   B3E.VOBE            1.
    BLV.3EBE  XRV.3EBE           12.   N13.3EBE           61.
    BLV.3EBE  N14.3EBE           68.   N24.3EBE           95.
    BLV.3EBE  N36.3EBE          100.   X15.3EBE           79.
    BLV.3EBE  WRO.3EBE      78.21025   WMO.3EBE      76.50646
    BN4.3EBE  BN4...BE           -1.   B3E.VOBE            1.
    BN4.3EBE  XRV.3EBE          61.3   N13.3EBE          145.
    BN4.3EBE  N14.3EBE          145.   N24.3EBE          100.
    BN4.3EBE  N36.3EBE          100.   X15.3EBE          125.
    BN4.3EBE  WRO.3EBE      94.25784   WMO.3EBE      88.35746
    BP8.3EBE  BP8...BE           -1.   B3E.VOBE            1.
    BP8.3EBE  XRV.3EBE            6.   N13.3EBE           3.6
    BP8.3EBE  N14.3EBE            4.   N24.3EBE          28.5
    BP8.3EBE  N36.3EBE           97.   X15.3EBE           6.2
    BP8.3EBE  WRO.3EBE     105.82863   WMO.3EBE      90.59047
    BAL.3PBE  BAL...BE           -1.   B3P.VOBE            1.
    BAL.3PBE  XRV.3PBE           8.2   N13.3PBE            9.
    BAL.3PBE  N14.3PBE           12.   N24.3PBE           80.
    BAL.3PBE  N36.3PBE           99.   X15.3PBE           16.
    BAL.3PBE  WRO.3PBE      93.61705   WMO.3PBE       88.6782
    BCC.3PBE  BCC...BE           -1.   B3P.VOBE            1.
    BCC.3PBE  XRV.3PBE          21.1   N13.3PBE         116.8
    BCC.3PBE  N14.3PBE          114.   N24.3PBE          100.
    BCC.3PBE  N36.3PBE          100.   X15.3PBE          100.
    BCC.3PBE  WRO.3PBE      99.22401   WMO.3PBE      80.06283
    BCH.3PBE  BCH...BE           -1.   B3P.VOBE            1.
    BCH.3PBE  XRV.3PBE            2.   N13.3PBE         -14.8
    BCH.3PBE  N14.3PBE          -12.   N24.3PBE           2.4
    BCH.3PBE  N36.3PBE           90.   X15.3PBE          -8.2
    BCH.3PBE  WRO.3PBE      83.80122   WMO.3PBE      74.69736
    BCL.3PBE  BCL...BE           -1.   B3P.VOBE            1.
    BCL.3PBE  XRV.3PBE            8.   N13.3PBE            2.
    BCL.3PBE  N14.3PBE            9.   N24.3PBE           95.
    BCL.3PBE  N36.3PBE          100.   X15.3PBE           43.
    BCL.3PBE  WRO.3PBE      90.17511   WMO.3PBE      79.19421
    BHH.3PBE  BHH...BE           -1.   B3P.VOBE            1.
    BHH.3PBE  XRV.3PBE            1.   N13.3PBE         -21.9
    BHH.3PBE  N14.3PBE          -18.   N24.3PBE            4.
    BHH.3PBE  N36.3PBE          96.5   X15.3PBE         -12.5
    BHH.3PBE  WRO.3PBE       85.9962   WMO.3PBE      78.56822
    BHL.3PBE  BHL...BE           -1.   B3P.VOBE            1.
    BHL.3PBE  XRV.3PBE          12.5   N13.3PBE           37.
    BHL.3PBE  N14.3PBE           49.   N24.3PBE           98.
    BHL.3PBE  N36.3PBE          100.   X15.3PBE           65.
    BHL.3PBE  WRO.3PBE      86.96338   WMO.3PBE      82.92224
    BHX.3PBE  BHX...BE           -1.   B3P.VOBE            1.
    BHX.3PBE  XRV.3PBE            1.   N13.3PBE         -21.9
    BHX.3PBE  N14.3PBE          -18.   N24.3PBE            4.
    BHX.3PBE  N36.3PBE          96.5   X15.3PBE          -12.
    BHX.3PBE  WRO.3PBE       93.1476   WMO.3PBE      82.59274
    BLV.3PBE  BLV...BE           -1.   B3P.VOBE            1.
    BLV.3PBE  XRV.3PBE           12.   N13.3PBE           61.
    BLV.3PBE  N14.3PBE           68.   N24.3PBE           95.
    BLV.3PBE  N36.3PBE          100.   X15.3PBE           79.
    BLV.3PBE  WRO.3PBE      78.21025   WMO.3PBE      76.50646
    BN4.3PBE  BN4...BE           -1.   B3P.VOBE            1.
    BN4.3PBE  XRV.3PBE          61.3   N13.3PBE          145.
    BN4.3PBE  N14.3PBE          145.   N24.3PBE          100.
    BN4.3PBE  N36.3PBE          100.   X15.3PBE          125.
    BN4.3PBE  WRO.3PBE      94.25784   WMO.3PBE      88.35746
    BP8.3PBE  BP8...BE           -1.   B3P.VOBE            1.
    BP8.3PBE  XRV.3PBE            6.   N13.3PBE           3.6
    BP8.3PBE  N14.3PBE            4.   N24.3PBE          28.5
    BP8.3PBE  N36.3PBE           97.   X15.3PBE           6.2
    BP8.3PBE  WRO.3PBE     105.82863   WMO.3PBE      90.59047
    BAL.3RBE  BAL...BE           -1.   B3R.VOBE            1.
    BAL.3RBE  XRV.3RBE           8.2   N13.3RBE            9.
    BAL.3RBE  N14.3RBE           12.   N24.3RBE           80.
    BAL.3RBE  N36.3RBE           99.   X15.3RBE           16.
    BAL.3RBE  WRO.3RBE      93.61705   WMO.3RBE       88.6782
    BCC.3RBE  BCC...BE           -1.   B3R.VOBE            1.
    BCC.3RBE  XRV.3RBE          21.1   N13.3RBE         116.8
    BCC.3RBE  N14.3RBE          114.   N24.3RBE          100.
    BCC.3RBE  N36.3RBE          100.   X15.3RBE          100.
    BCC.3RBE  WRO.3RBE      99.22401   WMO.3RBE      80.06283
    BCH.3RBE  BCH...BE           -1.   B3R.VOBE            1.
    BCH.3RBE  XRV.3RBE            2.   N13.3RBE         -14.8
    BCH.3RBE  N14.3RBE          -12.   N24.3RBE           2.4
    BCH.3RBE  N36.3RBE           90.   X15.3RBE          -8.2
    BCH.3RBE  WRO.3RBE      83.80122   WMO.3RBE      74.69736
    BCL.3RBE  BCL...BE           -1.   B3R.VOBE            1.
    BCL.3RBE  XRV.3RBE            8.   N13.3RBE            2.
    BCL.3RBE  N14.3RBE            9.   N24.3RBE           95.
    BCL.3RBE  N36.3RBE          100.   X15.3RBE           43.
    BCL.3RBE  WRO.3RBE      90.17511   WMO.3RBE      79.19421
    BHH.3RBE  BHH...BE           -1.   B3R.VOBE            1.
    BHH.3RBE  XRV.3RBE            1.   N13.3RBE         -21.9
    BHH.3RBE  N14.3RBE          -18.   N24.3RBE            4.
    BHH.3RBE  N36.3RBE          96.5   X15.3RBE         -12.5
    BHH.3RBE  WRO.3RBE       85.9962   WMO.3RBE      78.56822
    BHL.3RBE  BHL...BE           -1.   B3R.VOBE            1.
    BHL.3RBE  XRV.3RBE          12.5   N13.3RBE           37.
    BHL.3RBE  N14.3RBE           49.   N24.3RBE           98.
    BHL.3RBE  N36.3RBE          100.   X15.3RBE           65.
    BHL.3RBE  WRO.3RBE      86.96338   WMO.3RBE      82.92224
    BHX.3RBE  BHX...BE           -1.   B3R.VOBE            1.
    BHX.3RBE  XRV.3RBE            1.   N13.3RBE         -21.9
    BHX.3RBE  N14.3RBE          -18.   N24.3RBE            4.
    BHX.3RBE  N36.3RBE          96.5   X15.3RBE          -12.
    BHX.3RBE  WRO.3RBE       93.1476   WMO.3RBE      82.59274
    BLV.3RBE  BLV...BE           -1.   B3R.VOBE            1.
    BLV.3RBE  XRV.3RBE           12.   N13.3RBE           61.
    BLV.3RBE  N14.3RBE           68.   N24.3RBE           95.
    BLV.3RBE  N36.3RBE          100.   X15.3RBE           79.
    BLV.3RBE  WRO.3RBE      78.21025   WMO.3RBE      76.50646
    BN4.3RBE  BN4...BE           -1.   B3R.VOBE            1.
    BN4.3RBE  XRV.3RBE          61.3   N13.3RBE          145.
    BN4.3RBE  N14.3RBE          145.   N24.3RBE          100.
    BN4.3RBE  N36.3RBE          100.   X15.3RBE          125.
    BN4.3RBE  WRO.3RBE      94.25784   WMO.3RBE      88.35746
    BP8.3RBE  BP8...BE           -1.   B3R.VOBE            1.
    BP8.3RBE  XRV.3RBE            6.   N13.3RBE           3.6
    BP8.3RBE  N14.3RBE            4.   N24.3RBE          28.5
    BP8.3RBE  N36.3RBE           97.   X15.3RBE           6.2
    BP8.3RBE  WRO.3RBE     105.82863   WMO.3RBE      90.59047
    D3E...BE  B3E...BE           -1.
    D3P...BE  B3P...BE           -1.
    D3R...BE  B3R...BE           -1.
    D3T...BE  B3T...BE           -1.
    EAL...BE  BAL...BE            1.
    ECC...BE  BCC...BE            1.
    ECH...BE  BCH...BE            1.
    ECL...BE  BCL...BE            1.
    EHH...BE  BHH...BE            1.
    EHL...BE  BHL...BE            1.
    EHX...BE  BHX...BE            1.
    ELV...BE  BLV...BE            1.
    EN4...BE  BN4...BE            1.
    EP8...BE  BP8...BE            1.
    M3..3TBE  B3T...BE            1.   B3E...BE          -.12
    M3..3TBE  B3P...BE          -.38   B3R...BE           -.5
    J&,1IOBE  B&,1..BE            1.
    JAL1IOBE  BAL1..BE            1.   FAT...J.           -2.
    JCC1IOBE  BCC1..BE            1.   FAT...J.           -2.
    JCH1IOBE  BCH1..BE            1.   FAT...J.           -2.
    JCL1IOBE  BCL1..BE            1.   FAT...J.           -2.
    JHH1IOBE  BHH1..BE            1.   FAT...J.           -2.
    JHL1IOBE  BHL1..BE            1.   FAT...J.           -2.
    JHX1IOBE  BHX1..BE            1.   FAT...J.           -2.
    JLV1IOBE  BLV1..BE            1.   FAT...J.           -2.
    JN41IOBE  BN41..BE            1.
    JP81IOBE  BP81..BE            1.   FAT...J.           -2.
    J&,1MXBE  B&,1..BE           -1.   B&,2..BE            1.
    JAL1MXBE  FAT...J.          .002   BAL1..BE           -1.
    JAL1MXBE  BAL2..BE            1.
    JCC1MXBE  FAT...J.          .002   BCC1..BE           -1.
    JCC1MXBE  BCC2..BE            1.
    JCH1MXBE  FAT...J.          .002   BCH1..BE           -1.
    JCH1MXBE  BCH2..BE            1.
    JCL1MXBE  FAT...J.          .002   BCL1..BE           -1.
    JCL1MXBE  BCL2..BE            1.
    JHH1MXBE  FAT...J.          .002   BHH1..BE           -1.
    JHH1MXBE  BHH2..BE            1.
    JHL1MXBE  FAT...J.          .002   BHL1..BE           -1.
    JHL1MXBE  BHL2..BE            1.
    JHX1MXBE  FAT...J.          .001   BHX1..BE           -1.
    JHX1MXBE  BHX2..BE            1.
    JLV1MXBE  FAT...J.          .002   BLV1..BE           -1.
    JLV1MXBE  BLV2..BE            1.
    JN41MXBE  FAT...J.          .002   BN41..BE           -1.
    JN41MXBE  BN42..BE            1.
    JP81MXBE  FAT...J.          .002   BP81..BE           -1.
    JP81MXBE  BP82..BE            1.
    J&,1TGBE  B&,1..BE           -1.   B&,2..BE            1.
    JAL1TGBE  FAT...J.         -.002   BAL1..BE           -1.
    JAL1TGBE  BAL2..BE            1.
    JCC1TGBE  FAT...J.         -.002   BCC1..BE           -1.
    JCC1TGBE  BCC2..BE            1.
    JCH1TGBE  FAT...J.         -.002   BCH1..BE           -1.
    JCH1TGBE  BCH2..BE            1.
    JCL1TGBE  FAT...J.         -.002   BCL1..BE           -1.
    JCL1TGBE  BCL2..BE            1.
    JHH1TGBE  FAT...J.         -.002   BHH1..BE           -1.
    JHH1TGBE  BHH2..BE            1.
    JHL1TGBE  FAT...J.         -.002   BHL1..BE           -1.
    JHL1TGBE  BHL2..BE            1.
    JHX1TGBE  FAT...J.         -.001   BHX1..BE           -1.
    JHX1TGBE  BHX2..BE            1.
    JLV1TGBE  FAT...J.         -.002   BLV1..BE           -1.
    JLV1TGBE  BLV2..BE            1.
    JN41TGBE  FAT...J.         -.002   BN41..BE           -1.
    JN41TGBE  BN42..BE            1.
    JP81TGBE  FAT...J.         -.002   BP81..BE           -1.
    JP81TGBE  BP82..BE            1.
    QVO13EBE  B3E...BE            1.   B3E.VOBE           -1.
    QVO13EBE  XRV.3EBE          -8.7   N13.3EBE          -10.
    QVO13EBE  N14.3EBE          -10.   N24.3EBE          -50.
    QVO13EBE  N36.3EBE          -90.   X15.3EBE          -47.
    QVO13EBE  NOM.3EBE          -89.   NRO.3EBE          -93.
    QVO13PBE  B3P...BE            1.   B3P.VOBE           -1.
    QVO13PBE  XRV.3PBE          -8.7   N13.3PBE          -10.
    QVO13PBE  N14.3PBE          -10.   N24.3PBE          -50.
    QVO13PBE  N36.3PBE          -90.   X15.3PBE          -47.
    QVO13PBE  NOM.3PBE          -85.   NRO.3PBE          -89.
    QVO13RBE  B3R...BE            1.   B3R.VOBE           -1.
    QVO13RBE  XRV.3RBE          -8.7   N13.3RBE          -10.
    QVO13RBE  N14.3RBE          -10.   N24.3RBE          -50.
    QVO13RBE  N36.3RBE          -90.   X15.3RBE          -47.
    QVO13RBE  NOM.3RBE          -88.   NRO.3RBE          -91.
    WMO13EBE  WMO.3EBE           -1.   NOM.3EBE            .5
    WRO13EBE  WRO.3EBE           -1.   NOM.3EBE            .5
    WRO13EBE  NRO.3EBE            1.
    WMO13PBE  WMO.3PBE           -1.   NOM.3PBE            .5
    WRO13PBE  WRO.3PBE           -1.   NOM.3PBE            .5
    WRO13PBE  NRO.3PBE            1.
    WMO13RBE  WMO.3RBE           -1.   NOM.3RBE            .5
    WRO13RBE  WRO.3RBE           -1.   NOM.3RBE            .5
    WRO13RBE  NRO.3RBE            1.
    J&,2MXBE  B&,2..BE           -1.   B&,3..BE            1.
    JAL2MXBE  FAT...J.          .002   BAL2..BE           -1.
    JAL2MXBE  BAL3..BE            1.
    JCC2MXBE  FAT...J.          .002   BCC2..BE           -1.
    JCC2MXBE  BCC3..BE            1.
    JCH2MXBE  FAT...J.          .002   BCH2..BE           -1.
    JCH2MXBE  BCH3..BE            1.
    JCL2MXBE  FAT...J.          .002   BCL2..BE           -1.
    JCL2MXBE  BCL3..BE            1.
    JHH2MXBE  FAT...J.          .002   BHH2..BE           -1.
    JHH2MXBE  BHH3..BE            1.
    JHL2MXBE  FAT...J.          .002   BHL2..BE           -1.
    JHL2MXBE  BHL3..BE            1.
    JHX2MXBE  FAT...J.          .001   BHX2..BE           -1.
    JHX2MXBE  BHX3..BE            1.
    JLV2MXBE  FAT...J.          .002   BLV2..BE           -1.
    JLV2MXBE  BLV3..BE            1.
    JN42MXBE  FAT...J.          .002   BN42..BE           -1.
    JN42MXBE  BN43..BE            1.
    JP82MXBE  FAT...J.          .002   BP82..BE           -1.
    JP82MXBE  BP83..BE            1.
    J&,2TGBE  B&,2..BE           -1.   B&,3..BE            1.
    JAL2TGBE  FAT...J.         -.002   BAL2..BE           -1.
    JAL2TGBE  BAL3..BE            1.
    JCC2TGBE  FAT...J.         -.002   BCC2..BE           -1.
    JCC2TGBE  BCC3..BE            1.
    JCH2TGBE  FAT...J.         -.002   BCH2..BE           -1.
    JCH2TGBE  BCH3..BE            1.
    JCL2TGBE  FAT...J.         -.002   BCL2..BE           -1.
    JCL2TGBE  BCL3..BE            1.
    JHH2TGBE  FAT...J.         -.002   BHH2..BE           -1.
    JHH2TGBE  BHH3..BE            1.
    JHL2TGBE  FAT...J.         -.002   BHL2..BE           -1.
    JHL2TGBE  BHL3..BE            1.
    JHX2TGBE  FAT...J.         -.001   BHX2..BE           -1.
    JHX2TGBE  BHX3..BE            1.
    JLV2TGBE  FAT...J.         -.002   BLV2..BE           -1.
    JLV2TGBE  BLV3..BE            1.
    JN42TGBE  FAT...J.         -.002   BN42..BE           -1.
    JN42TGBE  BN43..BE            1.
    JP82TGBE  FAT...J.         -.002   BP82..BE           -1.
    JP82TGBE  BP83..BE            1.
    QVO23EBE  B3E...BE            1.   B3E.VOBE           -1.
    QVO23EBE  XRV.3EBE          -8.7   N13.3EBE          -10.
    QVO23EBE  N14.3EBE          -10.   N24.3EBE          -50.
    QVO23EBE  N36.3EBE          -90.   X15.3EBE          -47.
    QVO23EBE  NOM.3EBE          -89.   NRO.3EBE          -93.
    QVO23PBE  B3P...BE            1.   B3P.VOBE           -1.
    QVO23PBE  XRV.3PBE          -8.7   N13.3PBE          -10.
    QVO23PBE  N14.3PBE          -10.   N24.3PBE          -50.
    QVO23PBE  N36.3PBE          -90.   X15.3PBE          -47.
    QVO23PBE  NOM.3PBE          -85.   NRO.3PBE          -89.
    QVO23RBE  B3R...BE            1.   B3R.VOBE           -1.
    QVO23RBE  XRV.3RBE          -8.7   N13.3RBE          -10.
    QVO23RBE  N14.3RBE          -10.   N24.3RBE          -50.
    QVO23RBE  N36.3RBE          -90.   X15.3RBE          -47.
    QVO23RBE  NOM.3RBE          -88.   NRO.3RBE          -91.
    WMO23EBE  WMO.3EBE           -1.   NOM.3EBE            .5
    WRO23EBE  WRO.3EBE           -1.   NOM.3EBE            .5
    WRO23EBE  NRO.3EBE            1.
    WMO23PBE  WMO.3PBE           -1.   NOM.3PBE            .5
    WRO23PBE  WRO.3PBE           -1.   NOM.3PBE            .5
    WRO23PBE  NRO.3PBE            1.
    WMO23RBE  WMO.3RBE           -1.   NOM.3RBE            .5
    WRO23RBE  WRO.3RBE           -1.   NOM.3RBE            .5
    WRO23RBE  NRO.3RBE            1.
    J&,3MXBE  B&,3..BE           -1.   B&,4..BE            1.
    JAL3MXBE  FAT...J.          .002   BAL3..BE           -1.
    JAL3MXBE  BAL4..BE            1.
    JCC3MXBE  FAT...J.          .002   BCC3..BE           -1.
    JCC3MXBE  BCC4..BE            1.
    JCH3MXBE  FAT...J.          .002   BCH3..BE           -1.
    JCH3MXBE  BCH4..BE            1.
    JCL3MXBE  FAT...J.          .002   BCL3..BE           -1.
    JCL3MXBE  BCL4..BE            1.
    JHH3MXBE  FAT...J.          .002   BHH3..BE           -1.
    JHH3MXBE  BHH4..BE            1.
    JHL3MXBE  FAT...J.          .002   BHL3..BE           -1.
    JHL3MXBE  BHL4..BE            1.
    JHX3MXBE  FAT...J.          .001   BHX3..BE           -1.
    JHX3MXBE  BHX4..BE            1.
    JLV3MXBE  FAT...J.          .002   BLV3..BE           -1.
    JLV3MXBE  BLV4..BE            1.
    JN43MXBE  FAT...J.          .002   BN43..BE           -1.
    JN43MXBE  BN44..BE            1.
    JP83MXBE  FAT...J.          .002   BP83..BE           -1.
    JP83MXBE  BP84..BE            1.
    J&,3TGBE  B&,3..BE           -1.   B&,4..BE            1.
    JAL3TGBE  FAT...J.         -.002   BAL3..BE           -1.
    JAL3TGBE  BAL4..BE            1.
    JCC3TGBE  FAT...J.         -.002   BCC3..BE           -1.
    JCC3TGBE  BCC4..BE            1.
    JCH3TGBE  FAT...J.         -.002   BCH3..BE           -1.
    JCH3TGBE  BCH4..BE            1.
    JCL3TGBE  FAT...J.         -.002   BCL3..BE           -1.
    JCL3TGBE  BCL4..BE            1.
    JHH3TGBE  FAT...J.         -.002   BHH3..BE           -1.
    JHH3TGBE  BHH4..BE            1.
    JHL3TGBE  FAT...J.         -.002   BHL3..BE           -1.
    JHL3TGBE  BHL4..BE            1.
    JHX3TGBE  FAT...J.         -.001   BHX3..BE           -1.
    JHX3TGBE  BHX4..BE            1.
    JLV3TGBE  FAT...J.         -.002   BLV3..BE           -1.
    JLV3TGBE  BLV4..BE            1.
    JN43TGBE  FAT...J.         -.002   BN43..BE           -1.
    JN43TGBE  BN44..BE            1.
    JP83TGBE  FAT...J.         -.002   BP83..BE           -1.
    JP83TGBE  BP84..BE            1.
    QVO33EBE  B3E...BE            1.   B3E.VOBE           -1.
    QVO33EBE  XRV.3EBE          -8.7   N13.3EBE          -10.
    QVO33EBE  N14.3EBE          -10.   N24.3EBE          -50.
    QVO33EBE  N36.3EBE          -90.   X15.3EBE          -47.
    QVO33EBE  NOM.3EBE          -89.   NRO.3EBE          -93.
    QVO33PBE  B3P...BE            1.   B3P.VOBE           -1.
    QVO33PBE  XRV.3PBE          -8.7   N13.3PBE          -10.
    QVO33PBE  N14.3PBE          -10.   N24.3PBE          -50.
    QVO33PBE  N36.3PBE          -90.   X15.3PBE          -47.
    QVO33PBE  NOM.3PBE          -85.   NRO.3PBE          -89.
    QVO33RBE  B3R...BE            1.   B3R.VOBE           -1.
    QVO33RBE  XRV.3RBE          -8.7   N13.3RBE          -10.
    QVO33RBE  N14.3RBE          -10.   N24.3RBE          -50.
    QVO33RBE  N36.3RBE          -90.   X15.3RBE          -47.
    QVO33RBE  NOM.3RBE          -88.   NRO.3RBE          -91.
    WMO33EBE  WMO.3EBE           -1.   NOM.3EBE            .5
    WRO33EBE  WRO.3EBE           -1.   NOM.3EBE            .5
    WRO33EBE  NRO.3EBE            1.
    WMO33PBE  WMO.3PBE           -1.   NOM.3PBE            .5
    WRO33PBE  WRO.3PBE           -1.   NOM.3PBE            .5
    WRO33PBE  NRO.3PBE            1.
    WMO33RBE  WMO.3RBE           -1.   NOM.3RBE            .5
    WRO33RBE  WRO.3RBE           -1.   NOM.3RBE            .5
    WRO33RBE  NRO.3RBE            1.
    J&,4MXBE  B&,4..BE           -1.
    JAL4MXBE  FAT...J.            .1   BAL4..BE           -1.
    JCC4MXBE  FAT...J.            .1   BCC4..BE           -1.
    JCH4MXBE  FAT...J.            .1   BCH4..BE           -1.
    JCL4MXBE  FAT...J.            .1   BCL4..BE           -1.
    JHH4MXBE  FAT...J.            .1   BHH4..BE           -1.
    JHL4MXBE  FAT...J.            .1   BHL4..BE           -1.
    JHX4MXBE  FAT...J.            .1   BHX4..BE           -1.
    JLV4MXBE  FAT...J.            .1   BLV4..BE           -1.
    JN44MXBE  FAT...J.            .1   BN44..BE           -1.
    JP84MXBE  FAT...J.            .1   BP84..BE           -1.
    J&,4TGBE  B&,4..BE           -1.
    JAL4TGBE  FAT...J.           -.1   BAL4..BE           -1.
    JCC4TGBE  FAT...J.           -.1   BCC4..BE           -1.
    JCH4TGBE  FAT...J.           -.1   BCH4..BE           -1.
    JCL4TGBE  FAT...J.           -.1   BCL4..BE           -1.
    JHH4TGBE  FAT...J.           -.1   BHH4..BE           -1.
    JHL4TGBE  FAT...J.           -.1   BHL4..BE           -1.
    JHX4TGBE  FAT...J.           -.1   BHX4..BE           -1.
    JLV4TGBE  FAT...J.           -.1   BLV4..BE           -1.
    JN44TGBE  FAT...J.           -.1   BN44..BE           -1.
    JP84TGBE  FAT...J.           -.1   BP84..BE           -1.
    QVO43EBE  B3E...BE            1.   B3E.VOBE           -1.
    QVO43EBE  XRV.3EBE          -8.7   N13.3EBE          -10.
    QVO43EBE  N14.3EBE          -10.   N24.3EBE          -50.
    QVO43EBE  N36.3EBE          -90.   X15.3EBE          -47.
    QVO43EBE  NOM.3EBE          -89.   NRO.3EBE          -93.
    QVO43PBE  B3P...BE            1.   B3P.VOBE           -1.
    QVO43PBE  XRV.3PBE          -8.7   N13.3PBE          -10.
    QVO43PBE  N14.3PBE          -10.   N24.3PBE          -50.
    QVO43PBE  N36.3PBE          -90.   X15.3PBE          -47.
    QVO43PBE  NOM.3PBE          -85.   NRO.3PBE          -89.
    QVO43RBE  B3R...BE            1.   B3R.VOBE           -1.
    QVO43RBE  XRV.3RBE          -8.7   N13.3RBE          -10.
    QVO43RBE  N14.3RBE          -10.   N24.3RBE          -50.
    QVO43RBE  N36.3RBE          -90.   X15.3RBE          -47.
    QVO43RBE  NOM.3RBE          -88.   NRO.3RBE          -91.
    WMO43EBE  WMO.3EBE           -1.   NOM.3EBE            .5
    WRO43EBE  WRO.3EBE           -1.   NOM.3EBE            .5
    WRO43EBE  NRO.3EBE            1.
    WMO43PBE  WMO.3PBE           -1.   NOM.3PBE            .5
    WRO43PBE  WRO.3PBE           -1.   NOM.3PBE            .5
    WRO43PBE  NRO.3PBE            1.
    WMO43RBE  WMO.3RBE           -1.   NOM.3RBE            .5
    WRO43RBE  WRO.3RBE           -1.   NOM.3RBE            .5
    WRO43RBE  NRO.3RBE            1.
RHS
BOUNDS
 FX BOUND     J&,1IOBE            0.
 UP BOUND     JAL1IOBE           92.
 UP BOUND     JCC1IOBE           39.
 UP BOUND     JCH1IOBE           87.
 UP BOUND     JCL1IOBE           29.
 UP BOUND     JHH1IOBE            0.
 UP BOUND     JHL1IOBE           20.
 UP BOUND     JHX1IOBE            0.
 UP BOUND     JLV1IOBE           28.
 UP BOUND     JN41IOBE           20.
 UP BOUND     JP81IOBE           71.
 LO BOUND     J&,1MXBE            0.
 UP BOUND     JAL1MXBE          130.
 UP BOUND     JCC1MXBE           45.
 UP BOUND     JCH1MXBE           53.
 UP BOUND     JCL1MXBE           55.
 UP BOUND     JHH1MXBE           75.
 UP BOUND     JHL1MXBE          112.
 FX BOUND     JHX1MXBE            0.
 UP BOUND     JLV1MXBE           73.
 UP BOUND     JN41MXBE          480.
 UP BOUND     JP81MXBE          154.
 UP BOUND     J&,1TGBE          121.
 LO BOUND     JAL1TGBE           10.
 UP BOUND     JAL1TGBE           50.
 LO BOUND     JCC1TGBE            5.
 UP BOUND     JCC1TGBE           30.
 LO BOUND     JCH1TGBE           10.
 UP BOUND     JCH1TGBE           77.
 LO BOUND     JCL1TGBE            5.
 UP BOUND     JCL1TGBE           20.
 FX BOUND     JHH1TGBE            0.
 LO BOUND     JHL1TGBE           10.
 UP BOUND     JHL1TGBE           18.
 FX BOUND     JHX1TGBE            0.
 LO BOUND     JLV1TGBE            2.
 UP BOUND     JLV1TGBE            5.
 UP BOUND     JN41TGBE           20.
 LO BOUND     JP81TGBE           10.
 UP BOUND     JP81TGBE           71.
 LO BOUND     J&,2MXBE            0.
 UP BOUND     JAL2MXBE          130.
 UP BOUND     JCC2MXBE           55.
 UP BOUND     JCH2MXBE           93.
 UP BOUND     JCL2MXBE           60.
 UP BOUND     JHH2MXBE           75.
 UP BOUND     JHL2MXBE          115.
 FX BOUND     JHX2MXBE            0.
 UP BOUND     JLV2MXBE           67.
 UP BOUND     JN42MXBE          480.
 UP BOUND     JP82MXBE          154.
 UP BOUND     J&,2TGBE          121.
 LO BOUND     JAL2TGBE           10.
 UP BOUND     JAL2TGBE           50.
 LO BOUND     JCC2TGBE            5.
 UP BOUND     JCC2TGBE           20.
 LO BOUND     JCH2TGBE           10.
 UP BOUND     JCH2TGBE           37.
 LO BOUND     JCL2TGBE            5.
 UP BOUND     JCL2TGBE           15.
 FX BOUND     JHH2TGBE            0.
 LO BOUND     JHL2TGBE           10.
 UP BOUND     JHL2TGBE           15.
 FX BOUND     JHX2TGBE            0.
 LO BOUND     JLV2TGBE            5.
 UP BOUND     JLV2TGBE            8.
 UP BOUND     JN42TGBE           20.
 LO BOUND     JP82TGBE           10.
 UP BOUND     JP82TGBE           71.
 LO BOUND     J&,3MXBE            0.
 UP BOUND     JAL3MXBE          130.
 UP BOUND     JCC3MXBE           55.
 UP BOUND     JCH3MXBE           93.
 UP BOUND     JCL3MXBE           60.
 UP BOUND     JHH3MXBE           75.
 UP BOUND     JHL3MXBE          105.
 FX BOUND     JHX3MXBE            0.
 UP BOUND     JLV3MXBE           67.
 UP BOUND     JN43MXBE         4980.
 UP BOUND     JP83MXBE          154.
 UP BOUND     J&,3TGBE          110.
 LO BOUND     JAL3TGBE           10.
 UP BOUND     JAL3TGBE           50.
 LO BOUND     JCC3TGBE            5.
 UP BOUND     JCC3TGBE           20.
 LO BOUND     JCH3TGBE           10.
 UP BOUND     JCH3TGBE           37.
 LO BOUND     JCL3TGBE            5.
 UP BOUND     JCL3TGBE           15.
 FX BOUND     JHH3TGBE            0.
 LO BOUND     JHL3TGBE           10.
 UP BOUND     JHL3TGBE           25.
 FX BOUND     JHX3TGBE            0.
 LO BOUND     JLV3TGBE            5.
 UP BOUND     JLV3TGBE            8.
 UP BOUND     JN43TGBE           20.
 LO BOUND     JP83TGBE           10.
 UP BOUND     JP83TGBE           71.
 LO BOUND     J&,4MXBE            0.
 UP BOUND     JAL4MXBE           20.
 UP BOUND     JCC4MXBE           20.
 UP BOUND     JCH4MXBE           20.
 UP BOUND     JCL4MXBE           20.
 FX BOUND     JHH4MXBE            0.
 UP BOUND     JHL4MXBE           20.
 FX BOUND     JHX4MXBE            0.
 UP BOUND     JLV4MXBE           20.
 FX BOUND     JN44MXBE            0.
 UP BOUND     JP84MXBE           20.
 FX BOUND     J&,4TGBE            0.
 FX BOUND     JAL4TGBE            0.
 FX BOUND     JCC4TGBE            0.
 FX BOUND     JCH4TGBE            0.
 FX BOUND     JCL4TGBE            0.
 FX BOUND     JHH4TGBE            0.
 FX BOUND     JHL4TGBE            0.
 FX BOUND     JHX4TGBE            0.
 FX BOUND     JLV4TGBE            0.
 FX BOUND     JN44TGBE            0.
 FX BOUND     JP84TGBE            0.
ENDATA
