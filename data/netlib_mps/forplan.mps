* SCIP STATISTICS
*   Problem name     : FORPLAN
*   Variables        : 421 (0 binary, 0 integer, 0 implicit integer, 421 continuous)
*   Constraints      : 161
*   Obj. scale       : 1
*   Obj. offset      : 0
NAME          FORPLAN
OBJSENSE
  MIN
ROWS
 N  Obj 
 E  LC123 
 E  DEDO3_1R 
 E  DEDO3_2R 
 E  DEDO3_3R 
 E  DEDO3_4R 
 E  DEDO3_5R 
 E  DEDO3_6R 
 E  DEDO3_7R 
 E  DEDO3_8R 
 E  DEDO3_9R 
 E  DEDO310R 
 E  DEDO311R 
 E  DEDO312R 
 E  DEDO313R 
 E  DEDO314R 
 E  DEDO315R 
 E  DEDO5_1R 
 E  DEDO5_2R 
 E  DEDO5_3R 
 L  BR___1_1 
 L  BR___2_2 
 L  BR___2_3 
 E  VOLM_1_R 
 E  VOLM_2_R 
 E  VOLM_3_R 
 E  VOLM_4_R 
 E  VOLM_5_R 
 E  VOLM_6_R 
 E  VOLM_7_R 
 E  VOLM_8_R 
 E  VOLM_9_R 
 E  VOLM10_R 
 E  VOLM11_R 
 E  VOLM12_R 
 E  VOLM13_R 
 E  VOLM14_R 
 E  VOLM15_R 
 E  VOLM16_R 
 E  VOLM17_R 
 E  VOLM18_R 
 E  VOLM19_R 
 E  VOLM20_R 
 G  BHVG_2 
 L  BHVL_2 
 G  BHVG_3 
 L  BHVL_3 
 G  BHVG_4 
 L  BHVL_4 
 G  BHVG_5 
 L  BHVL_5 
 G  BHVG_6 
 L  BHVL_6 
 G  BHVG_7 
 L  BHVL_7 
 G  BHVG_8 
 L  BHVL_8 
 G  BHVG_9 
 L  BHVL_9 
 G  BHVG10 
 L  BHVL10 
 G  BHVG11 
 L  BHVL11 
 G  BHVG12 
 L  BHVL12 
 G  BHVG13 
 L  BHVL13 
 G  BHVG14 
 L  BHVL14 
 G  BHVG15 
 L  BHVL15 
 G  BHVG16 
 L  BHVL16 
 G  BHVG17 
 L  BHVL17 
 G  BHVG18 
 L  BHVL18 
 G  BHVG19 
 L  BHVL19 
 G  BHVG20 
 L  BHVL20 
 L  SYNDY 
 E  LTSY_R 
 L  LTSYCT 
 E  AVEINV_R 
 G  ENDINVCT 
 E  INVEN_R 
 L  A$___1_1 
 L  A$___1_2 
 L  A$_4-8_1 
 L  A$_4-8_2 
 L  A$_4-8_3 
 L  A$_4-8_4 
 E  GP+++_0R 
 L  GS+++_1R 
 L  GS+++_2R 
 L  GS+++_3R 
 L  GS+++_4R 
 L  GS+++_5R 
 L  GS+++_6R 
 L  GS+++_7R 
 L  GS+++_8R 
 L  GS+++_9R 
 L  GS+++10R 
 L  GS+++11R 
 L  GS+++12R 
 L  GS+++13R 
 L  GS+++14R 
 L  GS+++15R 
 E  GP---_0R 
 L  GS---_1R 
 L  GS---_2R 
 L  GS---_3R 
 L  GS---_4R 
 L  GS---_5R 
 L  GS---_6R 
 E  R012_MN1 
 E  R012_RD1 
 E  R012_TM1 
 E  R012_TM2 
 E  R012_TM3 
 E  R012_TM4 
 E  R012_TM5 
 E  R012_TM6 
 E  R012_TP1 
 E  R012_TP2 
 E  R012_TP3 
 E  R012_TP4 
 E  R012_TP5 
 E  R012_TP6 
 E  R037_MN1 
 E  R037_RD1 
 E  R037_TM2 
 E  R037_TP2 
 E  R048_MN1 
 E  R048_RD1 
 E  R048_TM1 
 E  R048_TM2 
 E  R048_TM3 
 E  R048_TM4 
 E  R048_TM5 
 E  R048_TP1 
 E  R048_TP2 
 E  R048_TP3 
 E  R048_TP4 
 E  R052_MN1 
 E  R052_RD1 
 E  R052_TM1 
 E  R052_TM2 
 E  R052_TM3 
 E  R052_TM4 
 E  R052_TM5 
 E  R083_MN1 
 E  R083_GM2 
 E  R083_RD1 
 E  R083_GR2 
 E  R092_MN2 
 E  R092_RD1 
 E  AZ__20 
 E  AZ__80 
 E  AZ__90 
 E  AZ_100 
COLUMNS
    DEDO3_11  Obj                        0.02466  DEDO3_1R                        -1 
    DEDO3_12  Obj                              0  DEDO3_1R                        -1 
    DEDO3_21  Obj                        0.01666  DEDO3_2R                        -1 
    DEDO3_22  Obj                              0  DEDO3_2R                        -1 
    DEDO3_31  Obj                        0.01125  DEDO3_3R                        -1 
    DEDO3_32  Obj                              0  DEDO3_3R                        -1 
    DEDO3_41  DEDO3_4R                        -1  Obj                         0.0076 
    DEDO3_42  DEDO3_4R                        -1  Obj                              0 
    DEDO3_51  DEDO3_5R                        -1  Obj                        0.00514 
    DEDO3_52  DEDO3_5R                        -1  Obj                              0 
    DEDO3_61  DEDO3_6R                        -1  Obj                        0.00347 
    DEDO3_62  Obj                              0  DEDO3_6R                        -1 
    DEDO3_71  Obj                        0.00234  DEDO3_7R                        -1 
    DEDO3_72  DEDO3_7R                        -1  Obj                              0 
    DEDO3_81  DEDO3_8R                        -1  Obj                        0.00158 
    DEDO3_82  Obj                              0  DEDO3_8R                        -1 
    DEDO3_91  Obj                        0.00107  DEDO3_9R                        -1 
    DEDO3_92  Obj                              0  DEDO3_9R                        -1 
    DEDO3101  Obj                        0.00072  DEDO310R                        -1 
    DEDO3102  Obj                              0  DEDO310R                        -1 
    DEDO3111  Obj                        0.00049  DEDO311R                        -1 
    DEDO3112  Obj                              0  DEDO311R                        -1 
    DEDO3121  Obj                        0.00033  DEDO312R                        -1 
    DEDO3122  DEDO312R                        -1  Obj                              0 
    DEDO3131  DEDO313R                        -1  Obj                        0.00022 
    DEDO3132  DEDO313R                        -1  Obj                              0 
    DEDO3141  DEDO314R                        -1  Obj                        0.00015 
    DEDO3142  DEDO314R                        -1  Obj                              0 
    DEDO3151  DEDO315R                        -1  Obj                         0.0001 
    DEDO3152  Obj                              0  DEDO315R                        -1 
    DEDO5_11  Obj                        0.12038  DEDO5_1R                        -1 
    DEDO5_12  DEDO5_1R                        -1  Obj                              0 
    DEDO5_21  Obj                        0.05019  DEDO5_2R                        -1 
    DEDO5_22  DEDO5_2R                        -1  Obj                              0 
    DEDO5_31  Obj                        0.00546  DEDO5_3R                        -1 
    DEDO5_32  Obj                              0  DEDO5_3R                        -1 
    VOLM_1    Obj                              0  VOLM_1_R                        -1 
    VOLM_1    BHVG_2                          -1 
    VOLM_2    Obj                              0  BHVG_2                           1 
    VOLM_2    VOLM_2_R                        -1  BHVG_3                          -1 
    VOLM_3    BHVG_3                           1  VOLM_3_R                        -1 
    VOLM_3    Obj                              0  BHVG_4                          -1 
    VOLM_4    Obj                              0  BHVG_5                          -1 
    VOLM_4    VOLM_4_R                        -1  BHVG_4                           1 
    VOLM_5    Obj                              0  VOLM_5_R                        -1 
    VOLM_5    BHVG_5                           1  BHVG_6                          -1 
    VOLM_6    BHVG_6                           1  BHVG_7                          -1 
    VOLM_6    Obj                              0  VOLM_6_R                        -1 
    VOLM_7    BHVG_7                           1  BHVG_8                          -1 
    VOLM_7    Obj                              0  VOLM_7_R                        -1 
    VOLM_8    BHVG_8                           1  VOLM_8_R                        -1 
    VOLM_8    BHVG_9                          -1  Obj                              0 
    VOLM_9    BHVG10                          -1  BHVG_9                           1 
    VOLM_9    Obj                              0  VOLM_9_R                        -1 
    VOLM10    Obj                              0  BHVG10                           1 
    VOLM10    BHVG11                          -1  VOLM10_R                        -1 
    VOLM11    BHVG11                           1  VOLM11_R                        -1 
    VOLM11    Obj                              0  BHVG12                          -1 
    VOLM12    BHVG12                           1  Obj                              0 
    VOLM12    BHVG13                          -1  VOLM12_R                        -1 
    VOLM13    BHVG13                           1  BHVG14                          -1 
    VOLM13    Obj                              0  VOLM13_R                        -1 
    VOLM14    BHVG15                          -1  Obj                              0 
    VOLM14    BHVG14                           1  VOLM14_R                        -1 
    VOLM15    Obj                              0  BHVG15                           1 
    VOLM15    BHVG16                          -1  VOLM15_R                        -1 
    VOLM16    BHVG16                           1  VOLM16_R                        -1 
    VOLM16    Obj                              0  BHVG17                          -1 
    VOLM17    BHVG17                           1  BHVG18                          -1 
    VOLM17    VOLM17_R                        -1  Obj                              0 
    VOLM18    BHVG19                          -1  Obj                              0 
    VOLM18    VOLM18_R                        -1  BHVG18                           1 
    VOLM19    BHVG19                           1  BHVG20                          -1 
    VOLM19    VOLM19_R                        -1  Obj                              0 
    VOLM20    BHVG20                           1  Obj                              0 
    VOLM20    VOLM20_R                        -1  SYNDY                            1 
    LTSY      SYNDY                           -1  LTSY_R                          -1 
    LTSY      Obj                              0  LTSYCT                           1 
    AVEINV    Obj                              0  AVEINV_R                        -1 
    AVEINV    ENDINVCT                        -1 
    INVEN     INVEN_R                         -1  ENDINVCT                         1 
    INVEN     Obj                              0 
    GP+++_0   GS+++_5R                     -0.18  GS+++_2R                     -0.18 
    GP+++_0   Obj                              0  GS+++_4R                     -0.18 
    GP+++_0   GS+++14R                     -0.18  GP+++_0R                        -1 
    GP+++_0   GS+++12R                     -0.18  GS+++11R                     -0.18 
    GP+++_0   GS+++15R                     -0.18  GS+++_1R                     -0.18 
    GP+++_0   GS+++13R                     -0.18  GS+++_8R                     -0.18 
    GP+++_0   GS+++_7R                     -0.18  GS+++10R                     -0.18 
    GP+++_0   GS+++_9R                     -0.18  GS+++_6R                     -0.18 
    GP+++_0   GS+++_3R                     -0.18 
    GP---_0   GS---_5R                    -0.012  GS---_6R                    -0.012 
    GP---_0   GP---_0R                        -1  GS---_2R                    -0.012 
    GP---_0   GS---_3R                    -0.012  GS---_4R                    -0.012 
    GP---_0   Obj                              0  GS---_1R                    -0.012 
    A___21_1  DEDO3_9R                   1.59091  R083_MN1                  -0.10606 
    A___21_1  DEDO3_8R                   1.59091  DEDO310R                   1.59091 
    A___21_1  DEDO3_1R                   1.59091  DEDO314R                   1.59091 
    A___21_1  R052_MN1                  -0.11742  DEDO3_7R                   1.59091 
    A___21_1  AZ__20                           1  DEDO3_6R                   1.59091 
    A___21_1  R012_MN1                  -0.37879  DEDO313R                   1.59091 
    A___21_1  Obj                              0  DEDO3_2R                   1.59091 
    A___21_1  DEDO3_5R                   1.59091  DEDO311R                   1.59091 
    A___21_1  DEDO3_3R                   1.59091  R048_MN1                  -0.24621 
    A___21_1  R037_MN1                  -0.15152  DEDO3_4R                   1.59091 
    A___21_1  DEDO315R                   1.59091  DEDO312R                   1.59091 
    A___22_1  DEDO312R                   2.46212  R083_RD1                  -0.10606 
    A___22_1  DEDO314R                   2.46212  R037_RD1                  -0.15152 
    A___22_1  DEDO3_9R                   2.46212  DEDO315R                   2.46212 
    A___22_1  R012_RD1                  -0.37879  DEDO3_3R                   2.46212 
    A___22_1  DEDO3_4R                   2.46212  Obj                      -0.022381 
    A___22_1  DEDO3_8R                   2.46212  R048_RD1                  -0.24621 
    A___22_1  DEDO310R                   2.46212  DEDO311R                   2.46212 
    A___22_1  R052_RD1                  -0.11742  DEDO3_1R                   2.02652 
    A___22_1  DEDO3_7R                   2.46212  DEDO3_2R                   2.46212 
    A___22_1  AZ__20                           1  DEDO3_5R                   2.46212 
    A___22_1  DEDO3_6R                   2.46212  DEDO313R                   2.46212 
    A___23_1  DEDO313R                   0.87121  DEDO3_5R                   0.87121 
    A___23_1  AZ__20                           1  R048_TM1                  -0.24621 
    A___23_1  DEDO5_3R                   2.95455  DEDO3_6R                   0.87121 
    A___23_1  DEDO3_1R                   1.23106  DEDO314R                   0.87121 
    A___23_1  DEDO3_7R                   0.87121  R012_TM1                  -0.37879 
    A___23_1  DEDO3_2R                   0.87121  DEDO310R                   0.87121 
    A___23_1  DEDO311R                   0.87121  DEDO5_1R                   2.46212 
    A___23_1  R052_TM1                  -0.11742  DEDO5_2R                   2.95455 
    A___23_1  DEDO315R                   0.87121  DEDO3_8R                   0.87121 
    A___23_1  DEDO3_4R                   0.87121  DEDO3_3R                   0.87121 
    A___23_1  Obj                         -0.314  R037_TM2                  -0.15152 
    A___23_1  DEDO3_9R                   0.87121  DEDO312R                   0.87121 
    A___23_1  R083_GR2                  -0.10606 
    A___23_2  R083_GR2                  -0.10606  DEDO5_2R                   2.95455 
    A___23_2  DEDO5_1R                   1.47727  DEDO3_1R                   1.59091 
    A___23_2  DEDO313R                   0.87121  DEDO5_3R                   2.95455 
    A___23_2  DEDO3_9R                   0.87121  DEDO3_2R                   1.23106 
    A___23_2  DEDO315R                   0.87121  R012_TM2                  -0.37879 
    A___23_2  DEDO312R                   0.87121  DEDO3_3R                   0.87121 
    A___23_2  R037_TM2                  -0.15152  DEDO3_4R                   0.87121 
    A___23_2  DEDO311R                   0.87121  Obj                        -0.2121 
    A___23_2  DEDO3_5R                   0.87121  R048_TM2                  -0.24621 
    A___23_2  AZ__20                           1  DEDO3_6R                   0.87121 
    A___23_2  DEDO3_8R                   0.87121  DEDO3_7R                   0.87121 
    A___23_2  DEDO310R                   0.87121  R052_TM2                  -0.11742 
    A___23_2  DEDO314R                   0.87121 
    A___81_1  DEDO314R                   1.32143  DEDO3_3R                   1.32143 
    A___81_1  DEDO3_1R                   1.32143  DEDO3_9R                   1.32143 
    A___81_1  DEDO310R                   1.32143  DEDO313R                   1.32143 
    A___81_1  DEDO3_4R                   1.32143  DEDO312R                   1.32143 
    A___81_1  DEDO311R                   1.32143  DEDO3_7R                   1.32143 
    A___81_1  R048_MN1                  -0.26786  Obj                              0 
    A___81_1  R083_MN1                  -0.20357  R092_MN2                  -0.06429 
    A___81_1  R037_MN1                  -0.14286  DEDO3_2R                   1.32143 
    A___81_1  R012_MN1                  -0.32143  DEDO3_5R                   1.32143 
    A___81_1  DEDO315R                   1.32143  AZ__80                           1 
    A___81_1  DEDO3_8R                   1.32143  DEDO3_6R                   1.32143 
    A___82_1  DEDO3_6R                   3.14286  R083_RD1                  -0.20357 
    A___82_1  DEDO3_8R                   3.14286  Obj                      -0.029358 
    A___82_1  AZ__80                           1  DEDO314R                   3.14286 
    A___82_1  DEDO3_9R                   3.14286  DEDO315R                   3.14286 
    A___82_1  DEDO3_1R                   2.23214  DEDO3_5R                   3.14286 
    A___82_1  DEDO3_3R                   3.14286  R037_RD1                  -0.14286 
    A___82_1  DEDO3_2R                   3.14286  R012_RD1                  -0.32143 
    A___82_1  DEDO311R                   3.14286  DEDO312R                   3.14286 
    A___82_1  DEDO3_4R                   3.14286  DEDO3_7R                   3.14286 
    A___82_1  DEDO313R                   3.14286  DEDO310R                   3.14286 
    A___82_1  R092_RD1                  -0.06429  R048_RD1                  -0.26786 
    A___83_1  DEDO310R                   0.71429  DEDO313R                   0.71429 
    A___83_1  Obj                       -0.35041  R092_MN2                  -0.06429 
    A___83_1  R048_TM1                  -0.26786  DEDO3_7R                   0.71429 
    A___83_1  DEDO3_4R                   0.71429  DEDO3_9R                   0.71429 
    A___83_1  DEDO315R                   0.71429  DEDO5_2R                      3.75 
    A___83_1  DEDO314R                   0.71429  AZ__80                           1 
    A___83_1  DEDO311R                   0.71429  DEDO312R                   0.71429 
    A___83_1  DEDO5_3R                      3.75  R012_TM1                  -0.32143 
    A___83_1  DEDO3_8R                   0.71429  DEDO3_1R                   1.03571 
    A___83_1  DEDO3_3R                   0.71429  DEDO3_2R                   0.71429 
    A___83_1  DEDO3_5R                   0.71429  DEDO5_1R                     3.125 
    A___83_1  R037_TM2                  -0.14286  R083_GR2                  -0.20357 
    A___83_1  DEDO3_6R                   0.71429 
    A___83_2  R037_TM2                  -0.14286  DEDO3_3R                   0.71429 
    A___83_2  Obj                       -0.23669  DEDO3_1R                   1.35714 
    A___83_2  DEDO311R                   0.71429  DEDO3_4R                   0.71429 
    A___83_2  DEDO313R                   0.71429  AZ__80                           1 
    A___83_2  DEDO315R                   0.71429  DEDO3_8R                   0.71429 
    A___83_2  DEDO3_9R                   0.71429  R048_TM2                  -0.26786 
    A___83_2  DEDO3_5R                   0.71429  DEDO310R                   0.71429 
    A___83_2  DEDO3_2R                   1.03571  DEDO3_6R                   0.71429 
    A___83_2  DEDO5_1R                     1.875  DEDO314R                   0.71429 
    A___83_2  DEDO5_2R                      3.75  DEDO312R                   0.71429 
    A___83_2  R083_GR2                  -0.20357  R012_TM2                  -0.32143 
    A___83_2  DEDO3_7R                   0.71429  R092_MN2                  -0.06429 
    A___83_2  DEDO5_3R                      3.75 
    A___84_1  DEDO5_3R                   3.21429  DEDO312R                   0.89286 
    A___84_1  DEDO3_2R                   0.89286  DEDO3_9R                   0.89286 
    A___84_1  DEDO315R                   0.89286  DEDO3_1R                     1.125 
    A___84_1  R012_TM1                  -0.23929  R083_GM2                  -0.20357 
    A___84_1  R012_TP1                  -0.08214  R037_TM2                  -0.05357 
    A___84_1  DEDO311R                   0.89286  R037_TP2                  -0.08929 
    A___84_1  DEDO3_8R                   0.89286  DEDO3_4R                   0.89286 
    A___84_1  DEDO310R                   0.89286  AZ__80                           1 
    A___84_1  R048_TM1                  -0.23214  DEDO3_5R                   0.89286 
    A___84_1  DEDO314R                   0.89286  DEDO3_6R                   0.89286 
    A___84_1  DEDO5_1R                   2.67857  Obj                       -0.31496 
    A___84_1  DEDO5_2R                   3.21429  R048_TP1                  -0.03571 
    A___84_1  DEDO3_7R                   0.89286  R092_MN2                  -0.06429 
    A___84_1  DEDO313R                   0.89286  DEDO3_3R                   0.89286 
    A___84_2  DEDO311R                   0.89286  Obj                       -0.21274 
    A___84_2  R037_TP2                  -0.08929  R012_TM2                  -0.23929 
    A___84_2  DEDO3_4R                   0.89286  R048_TM2                  -0.23214 
    A___84_2  DEDO3_5R                   0.89286  DEDO3_3R                   0.89286 
    A___84_2  R037_TM2                  -0.05357  R012_TP2                  -0.08214 
    A___84_2  DEDO314R                   0.89286  DEDO313R                   0.89286 
    A___84_2  DEDO3_6R                   0.89286  DEDO5_1R                   1.60714 
    A___84_2  DEDO3_2R                     1.125  DEDO310R                   0.89286 
    A___84_2  AZ__80                           1  DEDO5_2R                   3.21429 
    A___84_2  R048_TP2                  -0.03571  DEDO3_7R                   0.89286 
    A___84_2  R092_MN2                  -0.06429  DEDO5_3R                   3.21429 
    A___84_2  DEDO3_8R                   0.89286  DEDO3_1R                   1.35714 
    A___84_2  DEDO312R                   0.89286  R083_GM2                  -0.20357 
    A___84_2  DEDO315R                   0.89286  DEDO3_9R                   0.89286 
    A___91_1  Obj                              0  DEDO313R                   1.59091 
    A___91_1  DEDO314R                   1.59091  DEDO315R                   1.59091 
    A___91_1  R048_MN1                  -0.24621  DEDO3_6R                   1.59091 
    A___91_1  DEDO312R                   1.59091  R037_MN1                  -0.15152 
    A___91_1  DEDO311R                   1.59091  DEDO3_2R                   1.59091 
    A___91_1  DEDO3_9R                   1.59091  DEDO3_1R                   1.59091 
    A___91_1  DEDO310R                   1.59091  DEDO3_4R                   1.59091 
    A___91_1  DEDO3_8R                   1.59091  DEDO3_5R                   1.59091 
    A___91_1  R083_MN1                  -0.10606  DEDO3_7R                   1.59091 
    A___91_1  AZ__90                           1  R012_MN1                  -0.37879 
    A___91_1  DEDO3_3R                   1.59091  R052_MN1                  -0.11742 
    A___92_1  R052_RD1                  -0.11742  AZ__90                           1 
    A___92_1  DEDO3_3R                   2.46212  R048_RD1                  -0.24621 
    A___92_1  DEDO3_4R                   2.46212  DEDO3_7R                   2.46212 
    A___92_1  R012_RD1                  -0.37879  DEDO3_9R                   2.46212 
    A___92_1  DEDO3_5R                   2.46212  DEDO311R                   2.46212 
    A___92_1  R037_RD1                  -0.15152  DEDO3_6R                   2.46212 
    A___92_1  DEDO3_1R                   2.02652  DEDO3_8R                   2.46212 
    A___92_1  DEDO3_2R                   2.46212  DEDO312R                   2.46212 
    A___92_1  DEDO315R                   2.46212  DEDO313R                   2.46212 
    A___92_1  R083_RD1                  -0.10606  DEDO314R                   2.46212 
    A___92_1  DEDO310R                   2.46212  Obj                      -0.022381 
    A___93_1  DEDO314R                   0.87121  R048_TM3                  -0.06155 
    A___93_1  R048_TM4                  -0.06155  DEDO310R                   0.87121 
    A___93_1  DEDO5_1R                   2.46212  DEDO311R                   0.87121 
    A___93_1  DEDO3_2R                   0.87121  DEDO313R                   0.87121 
    A___93_1  DEDO3_6R                   0.87121  DEDO3_5R                   0.87121 
    A___93_1  R012_TM3                  -0.15909  R037_TM2                  -0.15152 
    A___93_1  DEDO315R                   0.87121  LC123                         2800 
    A___93_1  DEDO3_7R                   0.87121  DEDO5_3R                   2.95455 
    A___93_1  DEDO3_9R                   0.87121  DEDO3_3R                   0.87121 
    A___93_1  DEDO312R                   0.87121  R048_TM2                  -0.06155 
    A___93_1  AZ__90                           1  R048_TM1                  -0.06155 
    A___93_1  R052_TM1                  -0.02936  DEDO3_4R                   0.87121 
    A___93_1  R052_TM2                  -0.02936  R052_TM3                  -0.02936 
    A___93_1  R052_TM4                  -0.02936  DEDO3_1R                   1.23106 
    A___93_1  DEDO3_8R                   0.87121  Obj                         -0.314 
    A___93_1  R012_TM1                    -0.125  R083_GR2                  -0.10606 
    A___93_1  DEDO5_2R                   2.95455  R012_TM2                   -0.0947 
    A___93_2  DEDO313R                   0.87121  R012_TM2                    -0.125 
    A___93_2  DEDO5_3R                   2.95455  DEDO3_7R                   0.87121 
    A___93_2  DEDO5_2R                   2.95455  AZ__90                           1 
    A___93_2  DEDO3_9R                   0.87121  R083_GR2                  -0.10606 
    A___93_2  DEDO3_4R                   0.87121  R048_TM5                  -0.06155 
    A___93_2  R012_TM4                  -0.15909  DEDO3_8R                   0.87121 
    A___93_2  DEDO311R                   0.87121  R012_TM3                   -0.0947 
    A___93_2  LC123                         2800  DEDO3_1R                   1.59091 
    A___93_2  DEDO312R                   0.87121  DEDO3_6R                   0.87121 
    A___93_2  DEDO314R                   0.87121  DEDO3_2R                   1.23106 
    A___93_2  R048_TM4                  -0.06155  Obj                        -0.2121 
    A___93_2  DEDO310R                   0.87121  R048_TM3                  -0.06155 
    A___93_2  R052_TM2                  -0.02936  DEDO315R                   0.87121 
    A___93_2  R048_TM2                  -0.06155  R052_TM5                  -0.02936 
    A___93_2  DEDO3_3R                   0.87121  R052_TM4                  -0.02936 
    A___93_2  R037_TM2                  -0.15152  DEDO5_1R                   1.47727 
    A___93_2  DEDO3_5R                   0.87121  R052_TM3                  -0.02936 
    A__101_1  DEDO3_7R                   1.32143  DEDO3_5R                   1.32143 
    A__101_1  AZ_100                           1  R012_MN1                  -0.32143 
    A__101_1  DEDO3_9R                   1.32143  R037_MN1                  -0.14286 
    A__101_1  DEDO315R                   1.32143  DEDO3_3R                   1.32143 
    A__101_1  DEDO3_8R                   1.32143  DEDO3_4R                   1.32143 
    A__101_1  DEDO313R                   1.32143  Obj                              0 
    A__101_1  DEDO311R                   1.32143  DEDO312R                   1.32143 
    A__101_1  R092_MN2                  -0.06429  DEDO3_1R                   1.32143 
    A__101_1  DEDO3_6R                   1.32143  R083_MN1                  -0.20357 
    A__101_1  DEDO3_2R                   1.32143  R048_MN1                  -0.26786 
    A__101_1  DEDO310R                   1.32143  DEDO314R                   1.32143 
    A__102_1  R048_RD1                  -0.26786  DEDO311R                   3.14286 
    A__102_1  DEDO3_6R                   3.14286  R083_RD1                  -0.20357 
    A__102_1  DEDO313R                   3.14286  DEDO3_2R                   3.14286 
    A__102_1  DEDO310R                   3.14286  Obj                      -0.029358 
    A__102_1  DEDO3_1R                   2.23214  DEDO3_4R                   3.14286 
    A__102_1  DEDO3_3R                   3.14286  DEDO3_9R                   3.14286 
    A__102_1  R092_RD1                  -0.06429  DEDO3_8R                   3.14286 
    A__102_1  DEDO315R                   3.14286  DEDO3_7R                   3.14286 
    A__102_1  R037_RD1                  -0.14286  DEDO3_5R                   3.14286 
    A__102_1  R012_RD1                  -0.32143  DEDO312R                   3.14286 
    A__102_1  DEDO314R                   3.14286  AZ_100                           1 
    A__103_1  DEDO5_3R                      3.75  DEDO314R                   0.71429 
    A__103_1  DEDO312R                   0.71429  R083_GR2                  -0.20357 
    A__103_1  LC123                         2640  DEDO3_7R                   0.71429 
    A__103_1  R012_TM4                  -0.10607  DEDO5_2R                      3.75 
    A__103_1  DEDO3_9R                   0.71429  R092_MN2                  -0.06429 
    A__103_1  DEDO3_6R                   0.71429  R012_TM5                     -0.09 
    A__103_1  DEDO311R                   0.71429  R012_TM2                  -0.12536 
    A__103_1  R048_TM3                  -0.10446  DEDO3_2R                   0.71429 
    A__103_1  DEDO3_8R                   0.71429  DEDO3_5R                   0.71429 
    A__103_1  DEDO3_1R                   1.03571  R037_TM2                  -0.14286 
    A__103_1  DEDO5_1R                     3.125  R048_TM1                  -0.16339 
    A__103_1  DEDO3_3R                   0.71429  DEDO313R                   0.71429 
    A__103_1  Obj                       -0.35041  DEDO310R                   0.71429 
    A__103_1  DEDO3_4R                   0.71429  DEDO315R                   0.71429 
    A__103_1  AZ_100                           1 
    A__103_2  DEDO3_1R                   1.35714  DEDO3_8R                   0.71429 
    A__103_2  R037_TM2                  -0.14286  DEDO315R                   0.71429 
    A__103_2  DEDO5_1R                     1.875  Obj                       -0.23669 
    A__103_2  DEDO3_4R                   0.71429  DEDO310R                   0.71429 
    A__103_2  DEDO313R                   0.71429  DEDO3_3R                   0.71429 
    A__103_2  R012_TM5                  -0.10607  R048_TM2                  -0.16339 
    A__103_2  DEDO5_2R                      3.75  DEDO314R                   0.71429 
    A__103_2  DEDO3_5R                   0.71429  DEDO3_7R                   0.71429 
    A__103_2  DEDO312R                   0.71429  DEDO3_2R                   1.03571 
    A__103_2  R048_TM4                  -0.10446  LC123                         2640 
    A__103_2  R092_MN2                  -0.06429  DEDO3_6R                   0.71429 
    A__103_2  DEDO5_3R                      3.75  AZ_100                           1 
    A__103_2  R012_TM3                  -0.12536  DEDO3_9R                   0.71429 
    A__103_2  R083_GR2                  -0.20357  R012_TM6                     -0.09 
    A__103_2  DEDO311R                   0.71429 
    A__104_1  DEDO311R                   0.89286  DEDO3_9R                   0.89286 
    A__104_1  R048_TP1                  -0.01786  DEDO3_6R                   0.89286 
    A__104_1  R012_TM4                  -0.08854  DEDO312R                   0.89286 
    A__104_1  R092_MN2                  -0.06429  R012_TM2                  -0.09093 
    A__104_1  LC123                         2640  R048_TM3                  -0.09286 
    A__104_1  R048_TP3                  -0.01786  DEDO3_2R                   0.89286 
    A__104_1  R012_TP2                  -0.02136  DEDO3_5R                   0.89286 
    A__104_1  DEDO3_7R                   0.89286  R012_TM5                  -0.05982 
    A__104_1  R048_TM1                  -0.13929  DEDO314R                   0.89286 
    A__104_1  R083_GM2                  -0.20357  DEDO3_4R                   0.89286 
    A__104_1  DEDO5_2R                   3.21429  DEDO310R                   0.89286 
    A__104_1  Obj                       -0.31496  DEDO3_1R                     1.125 
    A__104_1  DEDO5_1R                   2.67857  R012_TP3                  -0.00739 
    A__104_1  DEDO3_3R                   0.89286  R012_TP5                  -0.03943 
    A__104_1  DEDO3_8R                   0.89286  R037_TM2                  -0.05357 
    A__104_1  DEDO313R                   0.89286  R037_TP2                  -0.08929 
    A__104_1  DEDO5_3R                   3.21429  AZ_100                           1 
    A__104_1  DEDO315R                   0.89286  R012_TP4                  -0.01396 
    A__104_2  DEDO5_1R                   1.60714  AZ_100                           1 
    A__104_2  DEDO315R                   0.89286  R012_TP3                  -0.02136 
    A__104_2  DEDO310R                   0.89286  R012_TP4                  -0.00739 
    A__104_2  LC123                         2640  R092_MN2                  -0.06429 
    A__104_2  DEDO5_2R                   3.21429  R012_TP5                  -0.01396 
    A__104_2  DEDO311R                   0.89286  DEDO3_1R                   1.35714 
    A__104_2  Obj                       -0.21274  DEDO314R                   0.89286 
    A__104_2  R012_TP6                  -0.03943  R012_TM6                  -0.05982 
    A__104_2  DEDO3_9R                   0.89286  DEDO5_3R                   3.21429 
    A__104_2  DEDO3_2R                     1.125  R012_TM5                  -0.08854 
    A__104_2  DEDO3_3R                   0.89286  R037_TM2                  -0.05357 
    A__104_2  R083_GM2                  -0.20357  R037_TP2                  -0.08929 
    A__104_2  DEDO313R                   0.89286  DEDO3_4R                   0.89286 
    A__104_2  DEDO3_8R                   0.89286  DEDO3_5R                   0.89286 
    A__104_2  R048_TM2                  -0.13929  DEDO3_7R                   0.89286 
    A__104_2  R048_TM4                  -0.09286  R048_TP4                  -0.01786 
    A__104_2  DEDO3_6R                   0.89286  R012_TM3                  -0.09093 
    A__104_2  DEDO312R                   0.89286  R048_TP2                  -0.01786 
    A__105_1  AZ_100                           1  DEDO312R                   0.89286 
    A__105_1  R048_TP1                  -0.01536  DEDO315R                   0.89286 
    A__105_1  R012_TP3                  -0.00739  DEDO310R                   0.89286 
    A__105_1  R012_TP4                  -0.01479  DEDO3_6R                   0.89286 
    A__105_1  R048_TP3                  -0.02036  R012_TM2                  -0.06461 
    A__105_1  R012_TP2                  -0.01643  R048_TM3                  -0.14161 
    A__105_1  R092_MN2                  -0.06429  DEDO3_7R                   0.89286 
    A__105_1  LC123                         2640  DEDO3_5R                   0.89286 
    A__105_1  DEDO5_2R                   3.21429  R048_TM1                  -0.09054 
    A__105_1  DEDO3_9R                   0.89286  DEDO3_4R                   0.89286 
    A__105_1  DEDO3_8R                   0.89286  R012_TP5                  -0.04354 
    A__105_1  DEDO311R                   0.89286  DEDO313R                   0.89286 
    A__105_1  DEDO3_1R                     1.125  R037_TP2                  -0.08929 
    A__105_1  R012_TM4                  -0.08375  Obj                       -0.29354 
    A__105_1  R037_TM2                  -0.05357  DEDO314R                   0.89286 
    A__105_1  R083_GM2                  -0.20357  DEDO5_3R                   3.21429 
    A__105_1  DEDO5_1R                   2.67857  DEDO3_3R                   0.89286 
    A__105_1  R012_TM5                  -0.09093  DEDO3_2R                   0.89286 
    A__105_2  DEDO5_3R                   3.21429  DEDO3_2R                     1.125 
    A__105_2  DEDO312R                   0.89286  R083_GM2                  -0.20357 
    A__105_2  DEDO3_3R                   0.89286  R012_TM5                  -0.08375 
    A__105_2  R012_TP6                  -0.04354  R012_TM6                  -0.09093 
    A__105_2  R012_TP5                  -0.01479  DEDO3_1R                   1.35714 
    A__105_2  AZ_100                           1  DEDO5_1R                   1.60714 
    A__105_2  DEDO314R                   0.89286  DEDO315R                   0.89286 
    A__105_2  Obj                       -0.19827  R037_TM2                  -0.05357 
    A__105_2  DEDO311R                   0.89286  DEDO313R                   0.89286 
    A__105_2  DEDO3_8R                   0.89286  R037_TP2                  -0.08929 
    A__105_2  DEDO3_4R                   0.89286  DEDO5_2R                   3.21429 
    A__105_2  DEDO3_7R                   0.89286  DEDO3_9R                   0.89286 
    A__105_2  LC123                         2640  DEDO3_5R                   0.89286 
    A__105_2  R048_TM2                  -0.09054  R048_TP4                  -0.02036 
    A__105_2  R012_TP4                  -0.00739  DEDO3_6R                   0.89286 
    A__105_2  R092_MN2                  -0.06429  R048_TP2                  -0.01536 
    A__105_2  R012_TM3                  -0.06461  R012_TP3                  -0.01643 
    A__105_2  DEDO310R                   0.89286  R048_TM4                  -0.14161 
    M012MN_1  Obj                     -0.0012632  R012_MN1                         1 
    M012RD_1  Obj                     -0.0010105  R012_RD1                         1 
    T012TM12  Obj                              0  R012_TM1                         1 
    T012TM12  R012_TM2                        -1 
    T012TM23  Obj                              0  R012_TM2                         1 
    T012TM23  R012_TM3                        -1 
    T012TM34  R012_TM3                         1  Obj                              0 
    T012TM34  R012_TM4                        -1 
    T012TM45  Obj                              0  R012_TM4                         1 
    T012TM45  R012_TM5                        -1 
    T012TM56  Obj                              0  R012_TM6                        -1 
    T012TM56  R012_TM5                         1 
    M012TF_1  AVEINV_R                   0.18843  Obj                        0.95137 
    M012TF_1  R012_TM1                         1  VOLM_1_R                     0.402 
    M012TF_1  INVEN_R                      0.267  A$___1_1                   0.01994 
    M012TF_1  VOLM_8_R                     0.361  GS+++_1R                         1 
    M012TF_1  LTSY_R                     0.05157  VOLM15_R                     0.361 
    M012TF_1  GS+++15R                         1  GP+++_0R                         1 
    M012TF_1  GS+++_8R                         1 
    M012TF_2  R012_TM1                         1  VOLM17_R                     0.367 
    M012TF_2  INVEN_R                      0.142  GP+++_0R                         1 
    M012TF_2  LTSY_R                     0.04587  Obj                          0.922 
    M012TF_2  GS+++_9R                         1  AVEINV_R                   0.21075 
    M012TF_2  A$___1_1                   0.01994  VOLM_9_R                     0.367 
    M012TF_2  GS+++_1R                         1  VOLM_1_R                     0.402 
    M012TF_3  VOLM_9_R                     0.361  GS+++_2R                         1 
    M012TF_3  AVEINV_R                   0.18843  LTSY_R                     0.05157 
    M012TF_3  GP+++_0R                         1  A$___1_2                   0.02448 
    M012TF_3  GS+++_9R                         1  VOLM16_R                     0.361 
    M012TF_3  Obj                        0.68813  INVEN_R                      0.203 
    M012TF_3  R012_TM2                         1  VOLM_2_R                     0.423 
    M012TF_4  VOLM_2_R                     0.423  Obj                        0.66835 
    M012TF_4  VOLM10_R                     0.367  LTSY_R                     0.04587 
    M012TF_4  VOLM18_R                     0.367  GS+++10R                         1 
    M012TF_4  GP+++_0R                         1  R012_TM2                         1 
    M012TF_4  AVEINV_R                   0.21075  A$___1_2                   0.02448 
    M012TF_4  INVEN_R                      0.022  GS+++_2R                         1 
    M012TF_5  AVEINV_R                   0.18843  VOLM_3_R                      0.44 
    M012TF_5  VOLM17_R                     0.361  GP+++_0R                         1 
    M012TF_5  Obj                        0.50325  GS+++_3R                         1 
    M012TF_5  VOLM10_R                     0.361  LTSY_R                     0.05157 
    M012TF_5  R012_TM3                         1  GS+++10R                         1 
    M012TF_5  INVEN_R                      0.142 
    M012TF_6  LTSY_R                     0.04587  AVEINV_R                   0.21075 
    M012TF_6  VOLM19_R                     0.367  GS+++_3R                         1 
    M012TF_6  R012_TM3                         1  GP+++_0R                         1 
    M012TF_6  Obj                        0.48975  VOLM_3_R                      0.44 
    M012TF_6  VOLM11_R                     0.367  GS+++11R                         1 
    M012TF_7  GS+++11R                         1  INVEN_R                      0.022 
    M012TF_7  Obj                         0.3852  VOLM11_R                     0.361 
    M012TF_7  GS+++_4R                         1  VOLM_4_R                     0.437 
    M012TF_7  LTSY_R                     0.05157  GP+++_0R                         1 
    M012TF_7  VOLM18_R                     0.361  AVEINV_R                   0.18843 
    M012TF_7  R012_TM4                         1 
    M012TF_8  VOLM_4_R                     0.437  GS+++12R                         1 
    M012TF_8  INVEN_R                      0.367  Obj                        0.37596 
    M012TF_8  R012_TM4                         1  VOLM20_R                     0.367 
    M012TF_8  VOLM12_R                     0.367  GS+++_4R                         1 
    M012TF_8  GP+++_0R                         1  AVEINV_R                   0.21075 
    M012TF_8  LTSY_R                     0.04587 
    M012TF_9  Obj                        0.26538  VOLM19_R                     0.361 
    M012TF_9  GS+++12R                         1  VOLM12_R                     0.361 
    M012TF_9  GP+++_0R                         1  R012_TM5                         1 
    M012TF_9  GS+++_5R                         1  VOLM_5_R                     0.429 
    M012TF_9  AVEINV_R                   0.18843  LTSY_R                     0.05157 
    M012TF_A  Obj                        0.25839  INVEN_R                      0.361 
    M012TF_A  GS+++13R                         1  R012_TM5                         1 
    M012TF_A  AVEINV_R                   0.21075  VOLM_5_R                     0.429 
    M012TF_A  VOLM13_R                     0.367  GP+++_0R                         1 
    M012TF_A  GS+++_5R                         1  LTSY_R                     0.04587 
    M012TF_B  GP+++_0R                         1  GS+++_6R                         1 
    M012TF_B  LTSY_R                     0.05157  INVEN_R                      0.361 
    M012TF_B  VOLM13_R                     0.361  VOLM20_R                     0.361 
    M012TF_B  GS+++13R                         1  R012_TM6                         1 
    M012TF_B  Obj                        0.17208  VOLM_6_R                     0.417 
    M012TF_B  AVEINV_R                   0.18843 
    M012TF_C  VOLM14_R                     0.367  LTSY_R                     0.04587 
    M012TF_C  GS+++_6R                         1  GP+++_0R                         1 
    M012TF_C  INVEN_R                      0.324  AVEINV_R                   0.21075 
    M012TF_C  Obj                        0.16728  GS+++14R                         1 
    M012TF_C  VOLM_6_R                     0.417  R012_TM6                         1 
    M012TF_D  R012_TM6                         1  Obj                        0.11232 
    M012TF_D  GS+++14R                         1  INVEN_R                      0.324 
    M012TF_D  GS+++_7R                         1  AVEINV_R                   0.18843 
    M012TF_D  VOLM_7_R                      0.41  LTSY_R                     0.05157 
    M012TF_D  GP+++_0R                         1  VOLM14_R                     0.361 
    M012TF_E  AVEINV_R                   0.21075  GP+++_0R                         1 
    M012TF_E  Obj                        0.10962  R012_TM6                         1 
    M012TF_E  VOLM15_R                     0.367  GS+++_7R                         1 
    M012TF_E  VOLM_7_R                      0.41  GS+++15R                         1 
    M012TF_E  LTSY_R                     0.04587  INVEN_R                      0.267 
    M012T1_1  VOLM12_R                     0.053  Obj                        0.98104 
    M012T1_1  VOLM_5_R                     0.053  GS+++15R                         1 
    M012T1_1  VOLM19_R                     0.053  LTSY_R                       0.056 
    M012T1_1  GS+++_1R                         1  INVEN_R                      0.244 
    M012T1_1  VOLM_8_R                     0.339  VOLM15_R                     0.339 
    M012T1_1  A$___1_1                   0.01994  R012_TM1                         1 
    M012T1_1  VOLM_1_R                     0.402  GP+++_0R                         1 
    M012T1_1  AVEINV_R                   0.18143  GS+++_8R                         1 
    M012T1_2  Obj                        0.95421  VOLM_5_R                     0.053 
    M012T1_2  LTSY_R                     0.04975  AVEINV_R                   0.20187 
    M012T1_2  GP+++_0R                         1  VOLM_9_R                     0.345 
    M012T1_2  GS+++_9R                         1  INVEN_R                      0.142 
    M012T1_2  VOLM13_R                     0.053  R012_TM1                         1 
    M012T1_2  VOLM17_R                     0.345  VOLM_1_R                     0.402 
    M012T1_2  A$___1_1                   0.01994  GS+++_1R                         1 
    M012T1_3  A$___1_1                   0.01994  VOLM_1_R                     0.402 
    M012T1_3  R012_TM1                         1  GP+++_0R                         1 
    M012T1_3  VOLM14_R                     0.053  AVEINV_R                   0.22133 
    M012T1_3  GS+++10R                         1  Obj                        0.93889 
    M012T1_3  VOLM_5_R                     0.053  GS+++_1R                         1 
    M012T1_3  VOLM10_R                     0.377  VOLM19_R                     0.377 
    M012T1_3  LTSY_R                     0.04778 
    M012T1_4  A$___1_2                   0.02448  AVEINV_R                   0.18143 
    M012T1_4  VOLM13_R                     0.053  INVEN_R                      0.206 
    M012T1_4  GS+++_9R                         1  GP+++_0R                         1 
    M012T1_4  Obj                         0.7082  VOLM20_R                     0.053 
    M012T1_4  R012_TM2                         1  VOLM_6_R                     0.053 
    M012T1_4  GS+++_2R                         1  VOLM_9_R                     0.339 
    M012T1_4  LTSY_R                       0.056  VOLM16_R                     0.339 
    M012T1_4  VOLM_2_R                     0.423 
    M012T1_5  VOLM_2_R                     0.423  GP+++_0R                         1 
    M012T1_5  INVEN_R                      0.022  VOLM10_R                     0.345 
    M012T1_5  GS+++_2R                         1  GS+++10R                         1 
    M012T1_5  Obj                        0.69013  VOLM18_R                     0.345 
    M012T1_5  VOLM14_R                     0.053  VOLM_6_R                     0.053 
    M012T1_5  LTSY_R                     0.04975  A$___1_2                   0.02448 
    M012T1_5  R012_TM2                         1  AVEINV_R                   0.20187 
    M012T1_6  R012_TM2                         1  GP+++_0R                         1 
    M012T1_6  AVEINV_R                   0.22133  INVEN_R                      0.377 
    M012T1_6  LTSY_R                     0.04778  Obj                        0.67971 
    M012T1_6  VOLM11_R                     0.377  VOLM_6_R                     0.053 
    M012T1_6  A$___1_2                   0.02448  VOLM20_R                     0.377 
    M012T1_6  VOLM15_R                     0.053  GS+++_2R                         1 
    M012T1_6  GS+++11R                         1  VOLM_2_R                     0.423 
    M012T1_7  AVEINV_R                   0.18143  VOLM_3_R                      0.44 
    M012T1_7  INVEN_R                      0.142  VOLM14_R                     0.053 
    M012T1_7  GP+++_0R                         1  LTSY_R                       0.056 
    M012T1_7  VOLM17_R                     0.339  R012_TM3                         1 
    M012T1_7  VOLM10_R                     0.339  Obj                        0.51675 
    M012T1_7  GS+++_3R                         1  GS+++10R                         1 
    M012T1_7  VOLM_7_R                     0.053 
    M012T1_8  VOLM_7_R                     0.053  LTSY_R                     0.04975 
    M012T1_8  VOLM19_R                     0.345  R012_TM3                         1 
    M012T1_8  GS+++_3R                         1  Obj                        0.50448 
    M012T1_8  VOLM11_R                     0.345  AVEINV_R                   0.20187 
    M012T1_8  VOLM15_R                     0.053  GS+++11R                         1 
    M012T1_8  GP+++_0R                         1  VOLM_3_R                      0.44 
    M012T1_9  GP+++_0R                         1  R012_TM3                         1 
    M012T1_9  VOLM12_R                     0.377  GS+++_3R                         1 
    M012T1_9  LTSY_R                     0.04778  INVEN_R                      0.345 
    M012T1_9  VOLM_7_R                     0.053  AVEINV_R                   0.22133 
    M012T1_9  GS+++12R                         1  Obj                        0.49638 
    M012T1_9  VOLM_3_R                      0.44  VOLM16_R                     0.053 
    M012T1_A  VOLM18_R                     0.339  AVEINV_R                   0.18143 
    M012T1_A  GP+++_0R                         1  R012_TM4                         1 
    M012T1_A  Obj                        0.39435  GS+++_4R                         1 
    M012T1_A  INVEN_R                      0.022  VOLM15_R                     0.053 
    M012T1_A  VOLM_8_R                     0.053  VOLM11_R                     0.339 
    M012T1_A  VOLM_4_R                     0.437  LTSY_R                       0.056 
    M012T1_A  GS+++11R                         1 
    M012T1_B  GS+++12R                         1  VOLM20_R                     0.345 
    M012T1_B  LTSY_R                     0.04975  VOLM12_R                     0.345 
    M012T1_B  VOLM_4_R                     0.437  Obj                        0.38592 
    M012T1_B  VOLM_8_R                     0.053  R012_TM4                         1 
    M012T1_B  GP+++_0R                         1  GS+++_4R                         1 
    M012T1_B  AVEINV_R                   0.20187  INVEN_R                      0.345 
    M012T1_B  VOLM16_R                     0.053 
    M012T1_C  GP+++_0R                         1  Obj                        0.38047 
    M012T1_C  VOLM_8_R                     0.053  AVEINV_R                   0.22133 
    M012T1_C  INVEN_R                      0.339  VOLM13_R                     0.377 
    M012T1_C  VOLM17_R                     0.053  GS+++13R                         1 
    M012T1_C  GS+++_4R                         1  LTSY_R                     0.04778 
    M012T1_C  R012_TM4                         1  VOLM_4_R                     0.437 
    M012T1_D  Obj                        0.27156  R012_TM5                         1 
    M012T1_D  VOLM_9_R                     0.053  VOLM12_R                     0.339 
    M012T1_D  GS+++12R                         1  VOLM_5_R                     0.429 
    M012T1_D  GP+++_0R                         1  AVEINV_R                   0.18143 
    M012T1_D  LTSY_R                       0.056  GS+++_5R                         1 
    M012T1_D  VOLM16_R                     0.053  VOLM19_R                     0.339 
    M012T1_E  INVEN_R                      0.339  VOLM17_R                     0.053 
    M012T1_E  AVEINV_R                   0.20187  GS+++_5R                         1 
    M012T1_E  LTSY_R                     0.04975  GP+++_0R                         1 
    M012T1_E  GS+++13R                         1  VOLM13_R                     0.345 
    M012T1_E  VOLM_5_R                     0.429  VOLM_9_R                     0.053 
    M012T1_E  Obj                        0.26512  R012_TM5                         1 
    M012T1_F  VOLM_5_R                     0.429  INVEN_R                      0.317 
    M012T1_F  GP+++_0R                         1  R012_TM5                         1 
    M012T1_F  LTSY_R                     0.04778  GS+++_5R                         1 
    M012T1_F  AVEINV_R                   0.22133  VOLM14_R                     0.377 
    M012T1_F  VOLM18_R                     0.053  Obj                        0.26223 
    M012T1_F  GS+++14R                         1  VOLM_9_R                     0.053 
    M012T1_G  Obj                        0.17627  GS+++13R                         1 
    M012T1_G  GS+++_6R                         1  AVEINV_R                   0.18143 
    M012T1_G  VOLM10_R                     0.053  GP+++_0R                         1 
    M012T1_G  INVEN_R                      0.339  VOLM20_R                     0.339 
    M012T1_G  R012_TM6                         1  VOLM13_R                     0.339 
    M012T1_G  LTSY_R                       0.056  VOLM17_R                     0.053 
    M012T1_G  VOLM_6_R                     0.417 
    M012T1_H  Obj                        0.17185  GS+++_6R                         1 
    M012T1_H  INVEN_R                      0.317  VOLM18_R                     0.053 
    M012T1_H  LTSY_R                     0.04975  VOLM14_R                     0.345 
    M012T1_H  AVEINV_R                   0.20187  VOLM10_R                     0.053 
    M012T1_H  GS+++14R                         1  R012_TM6                         1 
    M012T1_H  GP+++_0R                         1  VOLM_6_R                     0.417 
    M012T1_I  AVEINV_R                   0.22133  GP+++_0R                         1 
    M012T1_I  VOLM15_R                     0.377  R012_TM6                         1 
    M012T1_I  VOLM19_R                     0.053  LTSY_R                     0.04778 
    M012T1_I  VOLM10_R                     0.053  INVEN_R                      0.244 
    M012T1_I  Obj                         0.1697  GS+++_6R                         1 
    M012T1_I  GS+++15R                         1  VOLM_6_R                     0.417 
    M012T1_J  AVEINV_R                   0.18143  LTSY_R                       0.056 
    M012T1_J  R012_TM6                         1  GS+++_7R                         1 
    M012T1_J  Obj                        0.11518  GS+++14R                         1 
    M012T1_J  VOLM14_R                     0.339  VOLM_7_R                      0.41 
    M012T1_J  INVEN_R                      0.317  VOLM18_R                     0.053 
    M012T1_J  VOLM11_R                     0.053  GP+++_0R                         1 
    M012T1_K  Obj                        0.11271  VOLM11_R                     0.053 
    M012T1_K  GS+++15R                         1  GS+++_7R                         1 
    M012T1_K  INVEN_R                      0.244  VOLM_7_R                      0.41 
    M012T1_K  AVEINV_R                   0.20187  LTSY_R                     0.04975 
    M012T1_K  GP+++_0R                         1  VOLM15_R                     0.345 
    M012T1_K  R012_TM6                         1  VOLM19_R                     0.053 
    M012T1_L  VOLM11_R                     0.053  VOLM_7_R                      0.41 
    M012T1_L  GS+++_7R                         1  R012_TM6                         1 
    M012T1_L  GP+++_0R                         1  VOLM16_R                     0.377 
    M012T1_L  LTSY_R                     0.04778  VOLM20_R                     0.053 
    M012T1_L  Obj                        0.11145  INVEN_R                      0.206 
    M012T1_L  AVEINV_R                   0.22133 
    M012T1_M  VOLM_8_R                     0.356  R012_TM1                         1 
    M012T1_M  VOLM_1_R                     0.402  GS+++15R                         1 
    M012T1_M  INVEN_R                      0.242  LTSY_R                       0.059 
    M012T1_M  VOLM15_R                     0.356  VOLM20_R                     0.057 
    M012T1_M  GS+++_8R                         1  A$___1_1                   0.01994 
    M012T1_M  GS+++_1R                         1  GP+++_0R                         1 
    M012T1_M  VOLM_6_R                     0.057  AVEINV_R                   0.18343 
    M012T1_M  VOLM13_R                     0.057  Obj                        0.97761 
    M012T1_N  VOLM_1_R                     0.402  GS+++_9R                         1 
    M012T1_N  VOLM17_R                     0.362  A$___1_1                   0.01994 
    M012T1_N  LTSY_R                     0.05237  INVEN_R                      0.142 
    M012T1_N  VOLM_9_R                     0.362  AVEINV_R                   0.20575 
    M012T1_N  VOLM_6_R                     0.057  VOLM14_R                     0.057 
    M012T1_N  GP+++_0R                         1  Obj                        0.94944 
    M012T1_N  GS+++_1R                         1  R012_TM1                         1 
    M012T1_O  VOLM_1_R                     0.402  VOLM10_R                     0.388 
    M012T1_O  VOLM_6_R                     0.057  AVEINV_R                     0.226 
    M012T1_O  GS+++_1R                         1  GS+++10R                         1 
    M012T1_O  R012_TM1                         1  VOLM15_R                     0.057 
    M012T1_O  Obj                        0.93229  VOLM19_R                     0.388 
    M012T1_O  LTSY_R                     0.04944  GP+++_0R                         1 
    M012T1_O  A$___1_1                   0.01994 
    M012T1_P  INVEN_R                      0.203  GP+++_0R                         1 
    M012T1_P  R012_TM2                         1  Obj                        0.70581 
    M012T1_P  VOLM14_R                     0.057  VOLM_9_R                     0.356 
    M012T1_P  VOLM_2_R                     0.423  GS+++_2R                         1 
    M012T1_P  AVEINV_R                   0.18343  LTSY_R                       0.059 
    M012T1_P  VOLM_7_R                     0.057  VOLM16_R                     0.356 
    M012T1_P  GS+++_9R                         1  A$___1_2                   0.02448 
    M012T1_Q  A$___1_2                   0.02448  Obj                        0.68691 
    M012T1_Q  VOLM18_R                     0.362  INVEN_R                      0.022 
    M012T1_Q  GP+++_0R                         1  R012_TM2                         1 
    M012T1_Q  VOLM15_R                     0.057  VOLM_2_R                     0.423 
    M012T1_Q  VOLM10_R                     0.362  GS+++10R                         1 
    M012T1_Q  GS+++_2R                         1  VOLM_7_R                     0.057 
    M012T1_Q  LTSY_R                     0.05237  AVEINV_R                   0.20575 
    M012T1_R  GS+++11R                         1  Obj                        0.67525 
    M012T1_R  VOLM20_R                     0.388  AVEINV_R                     0.226 
    M012T1_R  A$___1_2                   0.02448  GS+++_2R                         1 
    M012T1_R  GP+++_0R                         1  VOLM16_R                     0.057 
    M012T1_R  VOLM11_R                     0.388  VOLM_7_R                     0.057 
    M012T1_R  LTSY_R                     0.04944  R012_TM2                         1 
    M012T1_R  INVEN_R                      0.388  VOLM_2_R                     0.423 
    M012T1_S  VOLM_3_R                      0.44  VOLM_8_R                     0.057 
    M012T1_S  VOLM15_R                     0.057  VOLM10_R                     0.356 
    M012T1_S  LTSY_R                       0.059  GS+++_3R                         1 
    M012T1_S  VOLM17_R                     0.356  INVEN_R                      0.142 
    M012T1_S  R012_TM3                         1  Obj                        0.51523 
    M012T1_S  AVEINV_R                   0.18343  GS+++10R                         1 
    M012T1_S  GP+++_0R                         1 
    M012T1_T  GP+++_0R                         1  GS+++_3R                         1 
    M012T1_T  AVEINV_R                   0.20575  VOLM11_R                     0.362 
    M012T1_T  R012_TM3                         1  VOLM16_R                     0.057 
    M012T1_T  VOLM19_R                     0.362  VOLM_3_R                      0.44 
    M012T1_T  Obj                        0.50231  LTSY_R                     0.05237 
    M012T1_T  GS+++11R                         1  VOLM_8_R                     0.057 
    M012T1_U  GS+++_3R                         1  AVEINV_R                     0.226 
    M012T1_U  VOLM12_R                     0.388  INVEN_R                      0.362 
    M012T1_U  VOLM17_R                     0.057  Obj                        0.49335 
    M012T1_U  GP+++_0R                         1  VOLM_3_R                      0.44 
    M012T1_U  R012_TM3                         1  VOLM_8_R                     0.057 
    M012T1_U  LTSY_R                     0.04944  GS+++12R                         1 
    M012T1_V  INVEN_R                      0.022  R012_TM4                         1 
    M012T1_V  GS+++_4R                         1  GS+++11R                         1 
    M012T1_V  VOLM16_R                     0.057  VOLM_9_R                     0.057 
    M012T1_V  GP+++_0R                         1  VOLM_4_R                     0.437 
    M012T1_V  VOLM18_R                     0.356  AVEINV_R                   0.18343 
    M012T1_V  LTSY_R                       0.059  VOLM11_R                     0.356 
    M012T1_V  Obj                        0.39333 
    M012T1_W  AVEINV_R                   0.20575  Obj                        0.38446 
    M012T1_W  VOLM20_R                     0.362  VOLM12_R                     0.362 
    M012T1_W  INVEN_R                      0.362  VOLM17_R                     0.057 
    M012T1_W  GS+++_4R                         1  VOLM_4_R                     0.437 
    M012T1_W  GP+++_0R                         1  LTSY_R                     0.05237 
    M012T1_W  R012_TM4                         1  GS+++12R                         1 
    M012T1_W  VOLM_9_R                     0.057 
    M012T1_X  AVEINV_R                     0.226  GP+++_0R                         1 
    M012T1_X  INVEN_R                      0.356  LTSY_R                     0.04944 
    M012T1_X  R012_TM4                         1  GS+++_4R                         1 
    M012T1_X  VOLM_4_R                     0.437  GS+++13R                         1 
    M012T1_X  Obj                        0.37843  VOLM18_R                     0.057 
    M012T1_X  VOLM_9_R                     0.057  VOLM13_R                     0.388 
    M012T1_Y  AVEINV_R                   0.18343  VOLM_5_R                     0.429 
    M012T1_Y  GS+++_5R                         1  GP+++_0R                         1 
    M012T1_Y  GS+++12R                         1  VOLM19_R                     0.356 
    M012T1_Y  R012_TM5                         1  LTSY_R                       0.059 
    M012T1_Y  Obj                        0.27088  VOLM12_R                     0.356 
    M012T1_Y  VOLM17_R                     0.057  VOLM10_R                     0.057 
    M012T1_Z  VOLM_5_R                     0.429  VOLM18_R                     0.057 
    M012T1_Z  GP+++_0R                         1  VOLM10_R                     0.057 
    M012T1_Z  LTSY_R                     0.05237  GS+++13R                         1 
    M012T1_Z  GS+++_5R                         1  R012_TM5                         1 
    M012T1_Z  INVEN_R                      0.356  VOLM13_R                     0.362 
    M012T1_Z  AVEINV_R                   0.20575  Obj                        0.26411 
    M012T1_[  VOLM19_R                     0.057  LTSY_R                     0.04944 
    M012T1_[  VOLM10_R                     0.057  INVEN_R                      0.319 
    M012T1_[  Obj                        0.26085  GP+++_0R                         1 
    M012T1_[  GS+++14R                         1  AVEINV_R                     0.226 
    M012T1_[  R012_TM5                         1  VOLM_5_R                     0.429 
    M012T1_[  VOLM14_R                     0.388  GS+++_5R                         1 
    M012T1_]  GS+++13R                         1  INVEN_R                      0.356 
    M012T1_]  VOLM_6_R                     0.417  AVEINV_R                   0.18343 
    M012T1_]  R012_TM6                         1  VOLM11_R                     0.057 
    M012T1_]  GS+++_6R                         1  Obj                        0.17581 
    M012T1_]  GP+++_0R                         1  VOLM20_R                     0.356 
    M012T1_]  VOLM13_R                     0.356  VOLM18_R                     0.057 
    M012T1_]  LTSY_R                       0.059 
    M012T1_#  GP+++_0R                         1  AVEINV_R                   0.20575 
    M012T1_#  GS+++14R                         1  VOLM11_R                     0.057 
    M012T1_#  R012_TM6                         1  INVEN_R                      0.319 
    M012T1_#  VOLM14_R                     0.362  VOLM19_R                     0.057 
    M012T1_#  LTSY_R                     0.05237  GS+++_6R                         1 
    M012T1_#  VOLM_6_R                     0.417  Obj                        0.17117 
    M012T1_^  GS+++15R                         1  GP+++_0R                         1 
    M012T1_^  VOLM_6_R                     0.417  VOLM11_R                     0.057 
    M012T1_^  INVEN_R                      0.242  R012_TM6                         1 
    M012T1_^  GS+++_6R                         1  LTSY_R                     0.04944 
    M012T1_^  VOLM15_R                     0.388  VOLM20_R                     0.057 
    M012T1_^  AVEINV_R                     0.226  Obj                        0.16877 
    M012T1_)  VOLM19_R                     0.057  VOLM14_R                     0.356 
    M012T1_)  INVEN_R                      0.319  VOLM12_R                     0.057 
    M012T1_)  Obj                        0.11485  GP+++_0R                         1 
    M012T1_)  GS+++_7R                         1  GS+++14R                         1 
    M012T1_)  R012_TM6                         1  AVEINV_R                   0.18343 
    M012T1_)  LTSY_R                       0.059  VOLM_7_R                      0.41 
    M012T1_-  Obj                        0.11225  INVEN_R                      0.242 
    M012T1_-  GS+++_7R                         1  LTSY_R                     0.05237 
    M012T1_-  AVEINV_R                   0.20575  GS+++15R                         1 
    M012T1_-  VOLM20_R                     0.057  VOLM15_R                     0.362 
    M012T1_-  VOLM12_R                     0.057  R012_TM6                         1 
    M012T1_-  GP+++_0R                         1  VOLM_7_R                      0.41 
    M012T1_+  VOLM16_R                     0.388  GP+++_0R                         1 
    M012T1_+  AVEINV_R                     0.226  R012_TM6                         1 
    M012T1_+  GS+++_7R                         1  Obj                        0.11073 
    M012T1_+  VOLM12_R                     0.057  LTSY_R                     0.04944 
    M012T1_+  VOLM_7_R                      0.41  INVEN_R                      0.203 
    M012T2_1  VOLM_5_R                      0.05  VOLM_7_R                     0.096 
    M012T2_1  R012_TM1                         1  Obj                         0.9872 
    M012T2_1  VOLM17_R                     0.343  GS+++_1R                         1 
    M012T2_1  GP+++_0R                         1  VOLM15_R                     0.096 
    M012T2_1  VOLM_1_R                     0.402  GS+++_9R                         1 
    M012T2_1  VOLM_9_R                     0.343  A$___1_1                   0.01994 
    M012T2_1  INVEN_R                      0.142  AVEINV_R                   0.20162 
    M012T2_1  VOLM13_R                      0.05  LTSY_R                     0.06112 
    M012T2_2  AVEINV_R                   0.22111  VOLM16_R                     0.096 
    M012T2_2  GP+++_0R                         1  A$___1_1                   0.01994 
    M012T2_2  VOLM10_R                     0.377  Obj                        0.97186 
    M012T2_2  VOLM19_R                     0.377  VOLM_5_R                      0.05 
    M012T2_2  R012_TM1                         1  LTSY_R                     0.05811 
    M012T2_2  VOLM14_R                      0.05  VOLM_1_R                     0.402 
    M012T2_2  VOLM_7_R                     0.096  GS+++_1R                         1 
    M012T2_2  GS+++10R                         1 
    M012T2_3  VOLM16_R                     0.096  VOLM_6_R                      0.05 
    M012T2_3  AVEINV_R                   0.20162  VOLM18_R                     0.343 
    M012T2_3  GS+++10R                         1  A$___1_2                   0.02448 
    M012T2_3  VOLM_2_R                     0.423  GP+++_0R                         1 
    M012T2_3  VOLM14_R                      0.05  LTSY_R                     0.06112 
    M012T2_3  Obj                        0.71248  INVEN_R                      0.022 
    M012T2_3  VOLM10_R                     0.343  VOLM_8_R                     0.096 
    M012T2_3  R012_TM2                         1  GS+++_2R                         1 
    M012T2_4  R012_TM2                         1  VOLM17_R                     0.096 
    M012T2_4  VOLM15_R                      0.05  VOLM11_R                     0.377 
    M012T2_4  VOLM_8_R                     0.096  LTSY_R                     0.05811 
    M012T2_4  AVEINV_R                   0.22111  A$___1_2                   0.02448 
    M012T2_4  Obj                        0.70202  GS+++11R                         1 
    M012T2_4  INVEN_R                      0.377  GP+++_0R                         1 
    M012T2_4  GS+++_2R                         1  VOLM_6_R                      0.05 
    M012T2_4  VOLM20_R                     0.377  VOLM_2_R                     0.423 
    M012T2_5  LTSY_R                     0.06112  VOLM_9_R                     0.096 
    M012T2_5  GS+++11R                         1  VOLM19_R                     0.343 
    M012T2_5  VOLM_3_R                      0.44  VOLM17_R                     0.096 
    M012T2_5  GP+++_0R                         1  VOLM11_R                     0.343 
    M012T2_5  AVEINV_R                   0.20162  VOLM_7_R                      0.05 
    M012T2_5  VOLM15_R                      0.05  GS+++_3R                         1 
    M012T2_5  R012_TM3                         1  Obj                        0.51961 
    M012T2_6  VOLM16_R                      0.05  VOLM_3_R                      0.44 
    M012T2_6  GS+++_3R                         1  R012_TM3                         1 
    M012T2_6  INVEN_R                      0.343  Obj                        0.51147 
    M012T2_6  LTSY_R                     0.05811  VOLM18_R                     0.096 
    M012T2_6  VOLM_9_R                     0.096  VOLM_7_R                      0.05 
    M012T2_6  GP+++_0R                         1  AVEINV_R                   0.22111 
    M012T2_6  VOLM12_R                     0.377  GS+++12R                         1 
    M012T2_7  VOLM12_R                     0.343  VOLM_8_R                      0.05 
    M012T2_7  GS+++12R                         1  VOLM_4_R                     0.437 
    M012T2_7  GP+++_0R                         1  VOLM18_R                     0.096 
    M012T2_7  AVEINV_R                   0.20162  VOLM10_R                     0.096 
    M012T2_7  GS+++_4R                         1  LTSY_R                     0.06112 
    M012T2_7  R012_TM4                         1  VOLM16_R                      0.05 
    M012T2_7  INVEN_R                      0.343  VOLM20_R                     0.343 
    M012T2_7  Obj                        0.39616 
    M012T2_8  VOLM19_R                     0.096  VOLM_8_R                      0.05 
    M012T2_8  INVEN_R                      0.339  GP+++_0R                         1 
    M012T2_8  VOLM10_R                     0.096  GS+++13R                         1 
    M012T2_8  VOLM17_R                      0.05  AVEINV_R                   0.22111 
    M012T2_8  VOLM_4_R                     0.437  GS+++_4R                         1 
    M012T2_8  R012_TM4                         1  LTSY_R                     0.05811 
    M012T2_8  VOLM13_R                     0.377  Obj                         0.3907 
    M012T2_9  VOLM_9_R                      0.05  VOLM19_R                     0.096 
    M012T2_9  AVEINV_R                   0.20162  VOLM13_R                     0.343 
    M012T2_9  LTSY_R                     0.06112  GS+++_5R                         1 
    M012T2_9  VOLM11_R                     0.096  Obj                        0.27209 
    M012T2_9  GP+++_0R                         1  VOLM17_R                      0.05 
    M012T2_9  R012_TM5                         1  GS+++13R                         1 
    M012T2_9  VOLM_5_R                     0.429  INVEN_R                      0.339 
    M012T2_A  VOLM_5_R                     0.429  VOLM11_R                     0.096 
    M012T2_A  LTSY_R                     0.05811  VOLM20_R                     0.096 
    M012T2_A  INVEN_R                      0.317  VOLM18_R                      0.05 
    M012T2_A  VOLM14_R                     0.377  GS+++_5R                         1 
    M012T2_A  VOLM_9_R                      0.05  GS+++14R                         1 
    M012T2_A  R012_TM5                         1  Obj                        0.26907 
    M012T2_A  GP+++_0R                         1  AVEINV_R                   0.22111 
    M012T2_B  Obj                        0.17649  GS+++14R                         1 
    M012T2_B  VOLM18_R                      0.05  VOLM20_R                     0.096 
    M012T2_B  VOLM12_R                     0.096  GP+++_0R                         1 
    M012T2_B  VOLM_6_R                     0.417  GS+++_6R                         1 
    M012T2_B  VOLM14_R                     0.343  INVEN_R                      0.317 
    M012T2_B  R012_TM6                         1  VOLM10_R                      0.05 
    M012T2_B  LTSY_R                     0.06112  AVEINV_R                   0.20162 
    M012T2_C  VOLM_6_R                     0.417  GS+++_6R                         1 
    M012T2_C  VOLM10_R                      0.05  VOLM15_R                     0.377 
    M012T2_C  R012_TM6                         1  GS+++15R                         1 
    M012T2_C  VOLM19_R                      0.05  Obj                        0.17422 
    M012T2_C  VOLM12_R                     0.096  GP+++_0R                         1 
    M012T2_C  AVEINV_R                   0.22111  LTSY_R                     0.05811 
    M012T2_C  INVEN_R                      0.244 
    M012T2_D  VOLM_7_R                      0.41  AVEINV_R                   0.20162 
    M012T2_D  VOLM19_R                      0.05  R012_TM6                         1 
    M012T2_D  GP+++_0R                         1  VOLM11_R                      0.05 
    M012T2_D  GS+++_7R                         1  INVEN_R                      0.244 
    M012T2_D  VOLM13_R                     0.096  LTSY_R                     0.06112 
    M012T2_D  Obj                        0.11574  GS+++15R                         1 
    M012T2_D  VOLM15_R                     0.343 
    M012T2_E  AVEINV_R                   0.22111  VOLM_7_R                      0.41 
    M012T2_E  VOLM16_R                     0.377  INVEN_R                      0.206 
    M012T2_E  Obj                        0.11453  LTSY_R                     0.05811 
    M012T2_E  R012_TM6                         1  GS+++_7R                         1 
    M012T2_E  VOLM13_R                     0.096  VOLM20_R                      0.05 
    M012T2_E  GP+++_0R                         1  VOLM11_R                      0.05 
    T012TP12  R012_TP1                         1  Obj                              0 
    T012TP12  R012_TP2                        -1 
    T012TP23  R012_TP3                        -1  R012_TP2                         1 
    T012TP23  Obj                              0 
    T012TP34  Obj                              0  R012_TP4                        -1 
    T012TP34  R012_TP3                         1 
    T012TP45  R012_TP5                        -1  Obj                              0 
    T012TP45  R012_TP4                         1 
    T012TP56  R012_TP5                         1  R012_TP6                        -1 
    T012TP56  Obj                              0 
    M012PF_1  VOLM15_R                   0.24548  VOLM_8_R                   0.24548 
    M012PF_1  VOLM_1_R                    0.2814  VOLM16_R                   0.11744 
    M012PF_1  INVEN_R                      0.267  Obj                         1.1767 
    M012PF_1  GS---_6R                      0.02  GS---_2R                      0.05 
    M012PF_1  GS---_1R                       0.1  AVEINV_R                   0.20521 
    M012PF_1  GS---_5R                   0.06667  VOLM_9_R                   0.11744 
    M012PF_1  VOLM_2_R                    0.1269  LTSY_R                     0.05185 
    M012PF_1  GP---_0R                         1  R012_TP1                         1 
    M012PF_2  INVEN_R                      0.142  VOLM_9_R                   0.24956 
    M012PF_2  LTSY_R                     0.04659  VOLM17_R                   0.24956 
    M012PF_2  VOLM_1_R                    0.2814  GS---_5R                   0.06667 
    M012PF_2  VOLM18_R                    0.1232  R012_TP1                         1 
    M012PF_2  GP---_0R                         1  GS---_2R                      0.05 
    M012PF_2  VOLM_2_R                    0.1269  AVEINV_R                   0.22615 
    M012PF_2  Obj                         1.1511  VOLM10_R                    0.1232 
    M012PF_2  GS---_6R                      0.02  GS---_1R                       0.1 
    M012PF_3  VOLM_9_R                   0.24548  INVEN_R                      0.203 
    M012PF_3  VOLM17_R                   0.11744  GS---_6R                      0.02 
    M012PF_3  LTSY_R                     0.05185  VOLM16_R                   0.24548 
    M012PF_3  VOLM_2_R                    0.2961  VOLM_3_R                     0.132 
    M012PF_3  R012_TP2                         1  GS---_5R                   0.06667 
    M012PF_3  Obj                        0.86847  GS---_2R                       0.1 
    M012PF_3  VOLM10_R                   0.11744  GP---_0R                         1 
    M012PF_3  AVEINV_R                   0.20521 
    M012PF_4  VOLM10_R                   0.24956  Obj                        0.85079 
    M012PF_4  VOLM_2_R                    0.2961  INVEN_R                      0.022 
    M012PF_4  GP---_0R                         1  R012_TP2                         1 
    M012PF_4  AVEINV_R                   0.22615  GS---_6R                      0.03 
    M012PF_4  VOLM18_R                   0.24956  VOLM11_R                    0.1232 
    M012PF_4  VOLM19_R                    0.1232  LTSY_R                     0.04659 
    M012PF_4  GS---_5R                   0.03333  GS---_2R                       0.1 
    M012PF_4  VOLM_3_R                     0.132 
    M012PF_5  GS---_5R                   0.03333  VOLM11_R                   0.11744 
    M012PF_5  GS---_3R                      0.05  GS---_6R                      0.03 
    M012PF_5  VOLM_3_R                     0.308  LTSY_R                     0.05185 
    M012PF_5  AVEINV_R                   0.20521  GS---_2R                      0.05 
    M012PF_5  VOLM18_R                   0.11744  VOLM_4_R                    0.1311 
    M012PF_5  GP---_0R                         1  VOLM17_R                   0.24548 
    M012PF_5  Obj                        0.63685  R012_TP3                         1 
    M012PF_5  INVEN_R                      0.142  VOLM10_R                   0.24548 
    M012PF_6  INVEN_R                     0.2541  VOLM12_R                    0.1232 
    M012PF_6  VOLM20_R                    0.1232  GS---_2R                      0.05 
    M012PF_6  GS---_6R                      0.04  GS---_3R                      0.05 
    M012PF_6  R012_TP3                         1  VOLM11_R                   0.24956 
    M012PF_6  Obj                        0.62437  VOLM_3_R                     0.308 
    M012PF_6  VOLM19_R                   0.24956  VOLM_4_R                    0.1311 
    M012PF_6  GP---_0R                         1  LTSY_R                     0.04659 
    M012PF_6  AVEINV_R                   0.22615 
    M012PF_7  VOLM11_R                   0.24548  AVEINV_R                   0.20521 
    M012PF_7  GS---_3R                       0.1  INVEN_R                      0.022 
    M012PF_7  VOLM18_R                   0.24548  VOLM12_R                   0.11744 
    M012PF_7  GS---_6R                      0.04  GP---_0R                         1 
    M012PF_7  VOLM_5_R                    0.1287  R012_TP4                         1 
    M012PF_7  Obj                         0.4633  VOLM_4_R                    0.3059 
    M012PF_7  LTSY_R                     0.05185  VOLM19_R                   0.11744 
    M012PF_8  R012_TP4                         1  LTSY_R                     0.04659 
    M012PF_8  Obj                        0.45475  VOLM_4_R                    0.3059 
    M012PF_8  GS---_6R                      0.03  GP---_0R                         1 
    M012PF_8  VOLM_5_R                    0.1287  VOLM12_R                   0.24956 
    M012PF_8  GS---_3R                       0.1  VOLM20_R                   0.24956 
    M012PF_8  INVEN_R                      0.367  VOLM13_R                    0.1232 
    M012PF_8  AVEINV_R                   0.22615 
    M012PF_9  VOLM12_R                   0.24548  INVEN_R                    0.24222 
    M012PF_9  VOLM_6_R                    0.1251  VOLM13_R                   0.11744 
    M012PF_9  GS---_3R                      0.05  AVEINV_R                   0.20521 
    M012PF_9  GS---_4R                      0.05  VOLM19_R                   0.24548 
    M012PF_9  GS---_6R                      0.04  Obj                        0.31453 
    M012PF_9  GP---_0R                         1  VOLM_5_R                    0.3003 
    M012PF_9  VOLM20_R                   0.11744  R012_TP5                         1 
    M012PF_9  LTSY_R                     0.05185 
    M012PF_A  AVEINV_R                   0.22615  VOLM_5_R                    0.3003 
    M012PF_A  INVEN_R                      0.361  VOLM13_R                   0.24956 
    M012PF_A  LTSY_R                     0.04659  GS---_3R                      0.05 
    M012PF_A  R012_TP5                         1  GP---_0R                         1 
    M012PF_A  Obj                        0.30826  GS---_4R                      0.05 
    M012PF_A  VOLM_6_R                    0.1251  GS---_6R                      0.02 
    M012PF_A  VOLM14_R                    0.1232 
    M012PF_B  INVEN_R                      0.361  VOLM14_R                   0.11744 
    M012PF_B  GS---_6R                      0.03  GS---_4R                       0.1 
    M012PF_B  AVEINV_R                   0.20521  R012_TP6                         1 
    M012PF_B  VOLM_6_R                    0.2919  Obj                        0.20637 
    M012PF_B  VOLM13_R                   0.24548  VOLM_7_R                     0.123 
    M012PF_B  GP---_0R                         1  LTSY_R                     0.05185 
    M012PF_B  VOLM20_R                   0.24548 
    M012PF_C  GS---_4R                       0.1  Obj                        0.20235 
    M012PF_C  R012_TP6                         1  GP---_0R                         1 
    M012PF_C  VOLM_6_R                    0.2919  VOLM_7_R                     0.123 
    M012PF_C  VOLM14_R                   0.24956  INVEN_R                      0.324 
    M012PF_C  AVEINV_R                   0.22615  LTSY_R                     0.04659 
    M012PF_C  VOLM15_R                    0.1232  GS---_6R                      0.02 
    M012PF_D  GS---_6R                      0.02  VOLM15_R                   0.11744 
    M012PF_D  AVEINV_R                   0.20521  R012_TP6                         1 
    M012PF_D  VOLM14_R                   0.24548  GS---_4R                      0.05 
    M012PF_D  VOLM_8_R                    0.1209  LTSY_R                     0.05185 
    M012PF_D  INVEN_R                      0.324  Obj                        0.13861 
    M012PF_D  VOLM_7_R                     0.287  GP---_0R                         1 
    M012PF_D  GS---_5R                   0.03333 
    M012PF_E  VOLM15_R                   0.24956  VOLM_8_R                    0.1209 
    M012PF_E  AVEINV_R                   0.22615  LTSY_R                     0.04659 
    M012PF_E  GS---_4R                      0.05  INVEN_R                      0.267 
    M012PF_E  R012_TP6                         1  Obj                         0.1362 
    M012PF_E  GS---_6R                      0.02  GS---_5R                   0.03333 
    M012PF_E  VOLM_7_R                     0.287  VOLM16_R                    0.1232 
    M012PF_E  GP---_0R                         1 
    M012P1_1  GP---_0R                         1  VOLM_2_R                    0.1269 
    M012P1_1  VOLM16_R                    0.1725  VOLM19_R                     0.053 
    M012P1_1  GS---_5R                   0.06667  VOLM_5_R                     0.053 
    M012P1_1  VOLM_1_R                    0.2814  Obj                         1.2023 
    M012P1_1  GS---_6R                      0.02  VOLM12_R                     0.053 
    M012P1_1  VOLM_9_R                    0.1725  LTSY_R                     0.05643 
    M012P1_1  GS---_2R                      0.05  AVEINV_R                   0.20607 
    M012P1_1  VOLM15_R                    0.1695  VOLM_8_R                    0.1695 
    M012P1_1  GS---_1R                       0.1  INVEN_R                      0.244 
    M012P1_1  R012_TP1                         1 
    M012P1_2  VOLM_5_R                     0.053  GS---_1R                       0.1 
    M012P1_2  INVEN_R                      0.142  AVEINV_R                   0.22544 
    M012P1_2  VOLM18_R                    0.1885  Obj                         1.1809 
    M012P1_2  GS---_2R                      0.05  LTSY_R                     0.05175 
    M012P1_2  VOLM_9_R                    0.1725  GS---_6R                      0.02 
    M012P1_2  R012_TP1                         1  VOLM13_R                     0.053 
    M012P1_2  VOLM_1_R                    0.2814  GS---_5R                   0.06667 
    M012P1_2  VOLM10_R                    0.1885  VOLM17_R                    0.1725 
    M012P1_2  VOLM_2_R                    0.1269  GP---_0R                         1 
    M012P1_3  VOLM_2_R                    0.1269  INVEN_R                    0.29325 
    M012P1_3  AVEINV_R                   0.24306  VOLM20_R                    0.1955 
    M012P1_3  VOLM_5_R                     0.053  GS---_2R                      0.05 
    M012P1_3  GS---_6R                      0.03  VOLM11_R                    0.1955 
    M012P1_3  R012_TP1                         1  VOLM_1_R                    0.2814 
    M012P1_3  VOLM10_R                    0.1885  VOLM14_R                     0.053 
    M012P1_3  GS---_1R                       0.1  GS---_5R                   0.03333 
    M012P1_3  GP---_0R                         1  Obj                         1.1645 
    M012P1_3  LTSY_R                     0.04856  VOLM19_R                    0.1885 
    M012P1_4  GP---_0R                         1  R012_TP2                         1 
    M012P1_4  VOLM_2_R                    0.2961  VOLM13_R                     0.053 
    M012P1_4  VOLM16_R                    0.1695  GS---_2R                       0.1 
    M012P1_4  GS---_6R                      0.02  INVEN_R                      0.206 
    M012P1_4  VOLM_6_R                     0.053  LTSY_R                     0.05643 
    M012P1_4  VOLM_3_R                     0.132  AVEINV_R                   0.20607 
    M012P1_4  GS---_5R                   0.06667  VOLM10_R                    0.1725 
    M012P1_4  VOLM17_R                    0.1725  Obj                        0.88581 
    M012P1_4  VOLM20_R                     0.053  VOLM_9_R                    0.1695 
    M012P1_5  VOLM19_R                    0.1885  VOLM11_R                    0.1885 
    M012P1_5  GP---_0R                         1  VOLM14_R                     0.053 
    M012P1_5  LTSY_R                     0.05175  R012_TP2                         1 
    M012P1_5  VOLM_6_R                     0.053  GS---_5R                   0.03333 
    M012P1_5  AVEINV_R                   0.22544  Obj                        0.87091 
    M012P1_5  VOLM_2_R                    0.2961  VOLM10_R                    0.1725 
    M012P1_5  INVEN_R                      0.022  GS---_2R                       0.1 
    M012P1_5  VOLM18_R                    0.1725  VOLM_3_R                     0.132 
    M012P1_5  GS---_6R                      0.03 
    M012P1_6  GS---_6R                      0.03  INVEN_R                      0.377 
    M012P1_6  AVEINV_R                   0.24306  VOLM12_R                    0.1955 
    M012P1_6  LTSY_R                     0.04856  Obj                        0.85957 
    M012P1_6  VOLM11_R                    0.1885  VOLM15_R                     0.053 
    M012P1_6  VOLM20_R                    0.1885  VOLM_2_R                    0.2961 
    M012P1_6  GP---_0R                         1  R012_TP2                         1 
    M012P1_6  VOLM_3_R                     0.132  VOLM_6_R                     0.053 
    M012P1_6  GS---_2R                       0.1 
    M012P1_7  VOLM_3_R                     0.308  VOLM14_R                     0.053 
    M012P1_7  GS---_5R                   0.03333  VOLM_4_R                    0.1311 
    M012P1_7  VOLM10_R                    0.1695  VOLM_7_R                     0.053 
    M012P1_7  GS---_3R                      0.05  GS---_6R                      0.03 
    M012P1_7  GS---_2R                      0.05  LTSY_R                     0.05643 
    M012P1_7  VOLM18_R                    0.1725  VOLM17_R                    0.1695 
    M012P1_7  AVEINV_R                   0.20607  Obj                        0.64853 
    M012P1_7  R012_TP3                         1  GP---_0R                         1 
    M012P1_7  VOLM11_R                    0.1725  INVEN_R                      0.142 
    M012P1_8  Obj                        0.63798  AVEINV_R                   0.22544 
    M012P1_8  VOLM15_R                     0.053  VOLM11_R                    0.1725 
    M012P1_8  VOLM20_R                    0.1885  VOLM19_R                    0.1725 
    M012P1_8  LTSY_R                     0.05175  GP---_0R                         1 
    M012P1_8  GS---_6R                      0.04  GS---_2R                      0.05 
    M012P1_8  GS---_3R                      0.05  VOLM_7_R                     0.053 
    M012P1_8  VOLM12_R                    0.1885  INVEN_R                    0.28275 
    M012P1_8  VOLM_3_R                     0.308  VOLM_4_R                    0.1311 
    M012P1_8  R012_TP3                         1 
    M012P1_9  VOLM_3_R                     0.308  R012_TP3                         1 
    M012P1_9  VOLM_4_R                    0.1311  GP---_0R                         1 
    M012P1_9  VOLM_7_R                     0.053  GS---_2R                      0.05 
    M012P1_9  GS---_6R                      0.02  VOLM12_R                    0.1885 
    M012P1_9  Obj                        0.62971  VOLM16_R                     0.053 
    M012P1_9  VOLM13_R                    0.1955  GS---_3R                      0.05 
    M012P1_9  INVEN_R                      0.345  LTSY_R                     0.04856 
    M012P1_9  AVEINV_R                   0.24306 
    M012P1_A  AVEINV_R                   0.20607  VOLM15_R                     0.053 
    M012P1_A  VOLM_8_R                     0.053  VOLM19_R                    0.1725 
    M012P1_A  Obj                         0.4712  GS---_3R                       0.1 
    M012P1_A  VOLM_5_R                    0.1287  R012_TP4                         1 
    M012P1_A  VOLM18_R                    0.1695  VOLM_4_R                    0.3059 
    M012P1_A  LTSY_R                     0.05643  VOLM12_R                    0.1725 
    M012P1_A  VOLM11_R                    0.1695  INVEN_R                      0.022 
    M012P1_A  GS---_6R                      0.04  GP---_0R                         1 
    M012P1_B  GP---_0R                         1  GS---_6R                      0.03 
    M012P1_B  VOLM12_R                    0.1725  VOLM_5_R                    0.1287 
    M012P1_B  VOLM20_R                    0.1725  GS---_3R                       0.1 
    M012P1_B  VOLM13_R                    0.1885  LTSY_R                     0.05175 
    M012P1_B  VOLM16_R                     0.053  R012_TP4                         1 
    M012P1_B  VOLM_4_R                    0.3059  VOLM_8_R                     0.053 
    M012P1_B  AVEINV_R                   0.22544  Obj                        0.46377 
    M012P1_B  INVEN_R                      0.345 
    M012P1_C  VOLM_4_R                    0.3059  INVEN_R                      0.339 
    M012P1_C  VOLM_8_R                     0.053  Obj                         0.4587 
    M012P1_C  AVEINV_R                   0.24306  R012_TP4                         1 
    M012P1_C  LTSY_R                     0.04856  VOLM13_R                    0.1885 
    M012P1_C  GS---_3R                       0.1  VOLM_5_R                    0.1287 
    M012P1_C  GS---_6R                      0.02  GP---_0R                         1 
    M012P1_C  VOLM14_R                    0.1955  VOLM17_R                     0.053 
    M012P1_D  AVEINV_R                   0.20607  VOLM_9_R                     0.053 
    M012P1_D  GS---_6R                      0.04  GS---_4R                      0.05 
    M012P1_D  VOLM12_R                    0.1695  INVEN_R                    0.25875 
    M012P1_D  VOLM19_R                    0.1695  GS---_3R                      0.05 
    M012P1_D  LTSY_R                     0.05643  VOLM20_R                    0.1725 
    M012P1_D  R012_TP5                         1  GP---_0R                         1 
    M012P1_D  VOLM13_R                    0.1725  VOLM16_R                     0.053 
    M012P1_D  VOLM_6_R                    0.1251  VOLM_5_R                    0.3003 
    M012P1_D  Obj                        0.31987 
    M012P1_E  VOLM13_R                    0.1725  VOLM_6_R                    0.1251 
    M012P1_E  VOLM17_R                     0.053  LTSY_R                     0.05175 
    M012P1_E  VOLM14_R                    0.1885  R012_TP5                         1 
    M012P1_E  GP---_0R                         1  GS---_3R                      0.05 
    M012P1_E  INVEN_R                      0.339  VOLM_5_R                    0.3003 
    M012P1_E  GS---_6R                      0.02  GS---_4R                      0.05 
    M012P1_E  VOLM_9_R                     0.053  AVEINV_R                   0.22544 
    M012P1_E  Obj                         0.3145 
    M012P1_F  VOLM18_R                     0.053  GS---_3R                      0.05 
    M012P1_F  INVEN_R                      0.317  VOLM_9_R                     0.053 
    M012P1_F  GS---_6R                      0.02  Obj                        0.31148 
    M012P1_F  R012_TP5                         1  VOLM_6_R                    0.1251 
    M012P1_F  GS---_4R                      0.05  LTSY_R                     0.04856 
    M012P1_F  VOLM_5_R                    0.3003  VOLM15_R                    0.1955 
    M012P1_F  GP---_0R                         1  AVEINV_R                   0.24306 
    M012P1_F  VOLM14_R                    0.1885 
    M012P1_G  GP---_0R                         1  VOLM14_R                    0.1725 
    M012P1_G  VOLM20_R                    0.1695  VOLM10_R                     0.053 
    M012P1_G  LTSY_R                     0.05643  VOLM13_R                    0.1695 
    M012P1_G  R012_TP6                         1  VOLM_6_R                    0.2919 
    M012P1_G  GS---_4R                       0.1  AVEINV_R                   0.20607 
    M012P1_G  GS---_6R                      0.03  Obj                        0.20984 
    M012P1_G  VOLM17_R                     0.053  VOLM_7_R                     0.123 
    M012P1_G  INVEN_R                      0.339 
    M012P1_H  INVEN_R                      0.317  Obj                         0.2066 
    M012P1_H  VOLM18_R                     0.053  VOLM_7_R                     0.123 
    M012P1_H  GS---_6R                      0.02  LTSY_R                     0.05175 
    M012P1_H  VOLM14_R                    0.1725  VOLM15_R                    0.1885 
    M012P1_H  GS---_4R                       0.1  R012_TP6                         1 
    M012P1_H  VOLM_6_R                    0.2919  AVEINV_R                   0.22544 
    M012P1_H  GP---_0R                         1  VOLM10_R                     0.053 
    M012P1_I  VOLM16_R                    0.1955  INVEN_R                      0.244 
    M012P1_I  VOLM15_R                    0.1885  VOLM19_R                     0.053 
    M012P1_I  AVEINV_R                   0.24306  VOLM_6_R                    0.2919 
    M012P1_I  GS---_6R                      0.02  R012_TP6                         1 
    M012P1_I  GS---_4R                       0.1  Obj                        0.20435 
    M012P1_I  VOLM10_R                     0.053  GP---_0R                         1 
    M012P1_I  LTSY_R                     0.04856  VOLM_7_R                     0.123 
    M012P1_J  GP---_0R                         1  VOLM14_R                    0.1695 
    M012P1_J  GS---_5R                   0.03333  VOLM_8_R                    0.1209 
    M012P1_J  VOLM_7_R                     0.287  AVEINV_R                   0.20607 
    M012P1_J  VOLM18_R                     0.053  VOLM11_R                     0.053 
    M012P1_J  Obj                        0.14112  LTSY_R                     0.05643 
    M012P1_J  INVEN_R                      0.317  R012_TP6                         1 
    M012P1_J  GS---_6R                      0.02  GS---_4R                      0.05 
    M012P1_J  VOLM15_R                    0.1725 
    M012P1_K  GS---_4R                      0.05  VOLM15_R                    0.1725 
    M012P1_K  GS---_6R                      0.02  R012_TP6                         1 
    M012P1_K  INVEN_R                      0.244  VOLM16_R                    0.1885 
    M012P1_K  GS---_5R                   0.03333  AVEINV_R                   0.22544 
    M012P1_K  LTSY_R                     0.05175  VOLM11_R                     0.053 
    M012P1_K  Obj                        0.13907  GP---_0R                         1 
    M012P1_K  VOLM_7_R                     0.287  VOLM_8_R                    0.1209 
    M012P1_K  VOLM19_R                     0.053 
    M012P1_L  AVEINV_R                   0.24306  VOLM11_R                     0.053 
    M012P1_L  VOLM_7_R                     0.287  VOLM_8_R                    0.1209 
    M012P1_L  GP---_0R                         1  Obj                        0.13772 
    M012P1_L  R012_TP6                         1  GS---_5R                   0.03333 
    M012P1_L  VOLM17_R                    0.1955  VOLM20_R                     0.053 
    M012P1_L  LTSY_R                     0.04856  GS---_4R                      0.05 
    M012P1_L  INVEN_R                      0.206  VOLM16_R                    0.1885 
    M012P1_L  GS---_6R                      0.02 
    M012P1_M  GS---_2R                      0.05  GS---_1R                       0.1 
    M012P1_M  VOLM_6_R                     0.057  VOLM20_R                     0.057 
    M012P1_M  VOLM_9_R                     0.181  AVEINV_R                   0.20929 
    M012P1_M  VOLM_2_R                    0.1269  GS---_6R                      0.02 
    M012P1_M  Obj                         1.1982  LTSY_R                     0.05943 
    M012P1_M  INVEN_R                      0.242  VOLM13_R                     0.057 
    M012P1_M  VOLM_8_R                     0.178  GS---_5R                   0.06667 
    M012P1_M  VOLM16_R                     0.181  VOLM_1_R                    0.2814 
    M012P1_M  R012_TP1                         1  VOLM15_R                     0.178 
    M012P1_M  GP---_0R                         1 
    M012P1_N  GP---_0R                         1  R012_TP1                         1 
    M012P1_N  VOLM_1_R                    0.2814  GS---_6R                      0.02 
    M012P1_N  AVEINV_R                      0.23  Obj                         1.1752 
    M012P1_N  VOLM_2_R                    0.1269  VOLM_6_R                     0.057 
    M012P1_N  GS---_1R                       0.1  GS---_5R                   0.06667 
    M012P1_N  VOLM17_R                     0.181  LTSY_R                       0.054 
    M012P1_N  VOLM_9_R                     0.181  VOLM18_R                     0.194 
    M012P1_N  INVEN_R                      0.142  GS---_2R                      0.05 
    M012P1_N  VOLM10_R                     0.194  VOLM14_R                     0.057 
    M012P1_O  LTSY_R                     0.05022  GS---_5R                   0.03333 
    M012P1_O  GS---_2R                      0.05  Obj                         1.1576 
    M012P1_O  GS---_1R                       0.1  VOLM15_R                     0.057 
    M012P1_O  VOLM11_R                     0.201  VOLM_6_R                     0.057 
    M012P1_O  GP---_0R                         1  VOLM19_R                     0.194 
    M012P1_O  VOLM_1_R                    0.2814  R012_TP1                         1 
    M012P1_O  GS---_6R                      0.03  AVEINV_R                   0.24833 
    M012P1_O  VOLM_2_R                    0.1269  VOLM10_R                     0.194 
    M012P1_O  INVEN_R                     0.3015  VOLM20_R                     0.201 
    M012P1_P  GS---_5R                   0.06667  LTSY_R                     0.05943 
    M012P1_P  INVEN_R                      0.203  VOLM14_R                     0.057 
    M012P1_P  VOLM17_R                     0.181  VOLM10_R                     0.181 
    M012P1_P  VOLM_9_R                     0.178  VOLM_2_R                    0.2961 
    M012P1_P  AVEINV_R                   0.20929  GS---_6R                      0.02 
    M012P1_P  VOLM16_R                     0.178  GP---_0R                         1 
    M012P1_P  VOLM_7_R                     0.057  R012_TP2                         1 
    M012P1_P  VOLM_3_R                     0.132  Obj                        0.88291 
    M012P1_P  GS---_2R                       0.1 
    M012P1_Q  INVEN_R                      0.022  AVEINV_R                      0.23 
    M012P1_Q  GP---_0R                         1  VOLM11_R                     0.194 
    M012P1_Q  VOLM_3_R                     0.132  GS---_5R                   0.03333 
    M012P1_Q  VOLM15_R                     0.057  VOLM18_R                     0.181 
    M012P1_Q  VOLM10_R                     0.181  VOLM_2_R                    0.2961 
    M012P1_Q  GS---_2R                       0.1  VOLM_7_R                     0.057 
    M012P1_Q  R012_TP2                         1  GS---_6R                      0.03 
    M012P1_Q  VOLM19_R                     0.194  LTSY_R                       0.054 
    M012P1_Q  Obj                        0.86704 
    M012P1_R  GP---_0R                         1  Obj                        0.85493 
    M012P1_R  GS---_2R                       0.1  VOLM12_R                     0.201 
    M012P1_R  VOLM_2_R                    0.2961  R012_TP2                         1 
    M012P1_R  VOLM16_R                     0.057  AVEINV_R                   0.24833 
    M012P1_R  VOLM_7_R                     0.057  GS---_6R                      0.03 
    M012P1_R  VOLM_3_R                     0.132  VOLM20_R                     0.194 
    M012P1_R  VOLM11_R                     0.194  LTSY_R                     0.05022 
    M012P1_R  INVEN_R                      0.388 
    M012P1_S  VOLM_8_R                     0.057  GS---_3R                      0.05 
    M012P1_S  VOLM15_R                     0.057  GS---_6R                      0.03 
    M012P1_S  VOLM_3_R                     0.308  VOLM11_R                     0.181 
    M012P1_S  LTSY_R                     0.05943  AVEINV_R                   0.20929 
    M012P1_S  INVEN_R                      0.142  GS---_5R                   0.03333 
    M012P1_S  GS---_2R                      0.05  VOLM_4_R                    0.1311 
    M012P1_S  VOLM10_R                     0.178  Obj                        0.64668 
    M012P1_S  VOLM18_R                     0.181  VOLM17_R                     0.178 
    M012P1_S  GP---_0R                         1  R012_TP3                         1 
    M012P1_T  LTSY_R                       0.054  GS---_3R                      0.05 
    M012P1_T  VOLM_8_R                     0.057  VOLM16_R                     0.057 
    M012P1_T  Obj                        0.63538  VOLM12_R                     0.194 
    M012P1_T  GS---_2R                      0.05  VOLM19_R                     0.181 
    M012P1_T  R012_TP3                         1  VOLM20_R                     0.194 
    M012P1_T  VOLM_4_R                    0.1311  INVEN_R                      0.291 
    M012P1_T  AVEINV_R                      0.23  VOLM_3_R                     0.308 
    M012P1_T  VOLM11_R                     0.181  GS---_6R                      0.04 
    M012P1_T  GP---_0R                         1 
    M012P1_U  INVEN_R                      0.362  VOLM_4_R                    0.1311 
    M012P1_U  LTSY_R                     0.05022  GP---_0R                         1 
    M012P1_U  GS---_2R                      0.05  VOLM_8_R                     0.057 
    M012P1_U  R012_TP3                         1  VOLM13_R                     0.201 
    M012P1_U  GS---_3R                      0.05  VOLM17_R                     0.057 
    M012P1_U  VOLM12_R                     0.194  AVEINV_R                   0.24833 
    M012P1_U  Obj                        0.62656  GS---_6R                      0.02 
    M012P1_U  VOLM_3_R                     0.308 
    M012P1_V  VOLM19_R                     0.181  VOLM_4_R                    0.3059 
    M012P1_V  GS---_6R                      0.04  VOLM11_R                     0.178 
    M012P1_V  Obj                        0.46995  R012_TP4                         1 
    M012P1_V  VOLM_9_R                     0.057  AVEINV_R                   0.20929 
    M012P1_V  INVEN_R                      0.022  GP---_0R                         1 
    M012P1_V  VOLM_5_R                    0.1287  GS---_3R                       0.1 
    M012P1_V  VOLM16_R                     0.057  VOLM18_R                     0.178 
    M012P1_V  VOLM12_R                     0.181  LTSY_R                     0.05943 
    M012P1_W  LTSY_R                       0.054  VOLM13_R                     0.194 
    M012P1_W  GP---_0R                         1  VOLM_5_R                    0.1287 
    M012P1_W  AVEINV_R                      0.23  INVEN_R                      0.362 
    M012P1_W  VOLM20_R                     0.181  VOLM17_R                     0.057 
    M012P1_W  GS---_3R                       0.1  VOLM_9_R                     0.057 
    M012P1_W  R012_TP4                         1  Obj                          0.462 
    M012P1_W  GS---_6R                      0.03  VOLM_4_R                    0.3059 
    M012P1_W  VOLM12_R                     0.181 
    M012P1_X  GS---_3R                       0.1  R012_TP4                         1 
    M012P1_X  INVEN_R                      0.356  GP---_0R                         1 
    M012P1_X  GS---_6R                      0.02  VOLM13_R                     0.194 
    M012P1_X  AVEINV_R                   0.24833  Obj                        0.45658 
    M012P1_X  VOLM_9_R                     0.057  LTSY_R                     0.05022 
    M012P1_X  VOLM14_R                     0.201  VOLM_5_R                    0.1287 
    M012P1_X  VOLM_4_R                    0.3059  VOLM18_R                     0.057 
    M012P1_Y  VOLM_5_R                    0.3003  GS---_3R                      0.05 
    M012P1_Y  VOLM20_R                     0.181  LTSY_R                     0.05943 
    M012P1_Y  VOLM17_R                     0.057  VOLM10_R                     0.057 
    M012P1_Y  Obj                        0.31903  VOLM19_R                     0.178 
    M012P1_Y  VOLM_6_R                    0.1251  VOLM13_R                     0.181 
    M012P1_Y  AVEINV_R                   0.20929  R012_TP5                         1 
    M012P1_Y  GS---_6R                      0.04  GP---_0R                         1 
    M012P1_Y  GS---_4R                      0.05  INVEN_R                     0.2715 
    M012P1_Y  VOLM12_R                     0.178 
    M012P1_Z  Obj                         0.3133  INVEN_R                      0.356 
    M012P1_Z  LTSY_R                       0.054  AVEINV_R                      0.23 
    M012P1_Z  VOLM18_R                     0.057  GS---_3R                      0.05 
    M012P1_Z  GP---_0R                         1  GS---_4R                      0.05 
    M012P1_Z  VOLM14_R                     0.194  VOLM_6_R                    0.1251 
    M012P1_Z  GS---_6R                      0.02  VOLM_5_R                    0.3003 
    M012P1_Z  R012_TP5                         1  VOLM10_R                     0.057 
    M012P1_Z  VOLM13_R                     0.181 
    M012P1_[  VOLM19_R                     0.057  R012_TP5                         1 
    M012P1_[  GS---_4R                      0.05  VOLM_5_R                    0.3003 
    M012P1_[  GS---_6R                      0.02  VOLM_6_R                    0.1251 
    M012P1_[  VOLM14_R                     0.194  LTSY_R                     0.05022 
    M012P1_[  VOLM15_R                     0.201  GS---_3R                      0.05 
    M012P1_[  GP---_0R                         1  VOLM10_R                     0.057 
    M012P1_[  AVEINV_R                   0.24833  INVEN_R                      0.319 
    M012P1_[  Obj                        0.31005 
    M012P1_]  VOLM18_R                     0.057  VOLM14_R                     0.181 
    M012P1_]  LTSY_R                     0.05943  VOLM_7_R                     0.123 
    M012P1_]  VOLM13_R                     0.178  GS---_4R                       0.1 
    M012P1_]  GS---_6R                      0.03  VOLM_6_R                    0.2919 
    M012P1_]  GP---_0R                         1  R012_TP6                         1 
    M012P1_]  VOLM11_R                     0.057  VOLM20_R                     0.178 
    M012P1_]  INVEN_R                      0.356  AVEINV_R                   0.20929 
    M012P1_]  Obj                        0.20926 
    M012P1_#  GS---_6R                      0.02  VOLM_6_R                    0.2919 
    M012P1_#  GS---_4R                       0.1  VOLM_7_R                     0.123 
    M012P1_#  VOLM14_R                     0.181  VOLM11_R                     0.057 
    M012P1_#  GP---_0R                         1  INVEN_R                      0.319 
    M012P1_#  VOLM15_R                     0.194  Obj                        0.20578 
    M012P1_#  LTSY_R                       0.054  VOLM19_R                     0.057 
    M012P1_#  AVEINV_R                      0.23  R012_TP6                         1 
    M012P1_^  Obj                        0.20339  GS---_4R                       0.1 
    M012P1_^  VOLM16_R                     0.201  R012_TP6                         1 
    M012P1_^  GP---_0R                         1  AVEINV_R                   0.24833 
    M012P1_^  LTSY_R                     0.05022  VOLM20_R                     0.057 
    M012P1_^  VOLM_7_R                     0.123  VOLM11_R                     0.057 
    M012P1_^  GS---_6R                      0.02  INVEN_R                      0.242 
    M012P1_^  VOLM15_R                     0.194  VOLM_6_R                    0.2919 
    M012P1_)  VOLM15_R                     0.181  VOLM12_R                     0.057 
    M012P1_)  AVEINV_R                   0.20929  VOLM_8_R                    0.1209 
    M012P1_)  Obj                        0.14072  VOLM_7_R                     0.287 
    M012P1_)  GP---_0R                         1  INVEN_R                      0.319 
    M012P1_)  GS---_5R                   0.03333  R012_TP6                         1 
    M012P1_)  VOLM14_R                     0.178  LTSY_R                     0.05943 
    M012P1_)  GS---_4R                      0.05  VOLM19_R                     0.057 
    M012P1_)  GS---_6R                      0.02 
    M012P1_-  GS---_4R                      0.05  GS---_5R                   0.03333 
    M012P1_-  INVEN_R                      0.242  VOLM_7_R                     0.287 
    M012P1_-  AVEINV_R                      0.23  GS---_6R                      0.02 
    M012P1_-  VOLM15_R                     0.181  Obj                        0.13852 
    M012P1_-  VOLM_8_R                    0.1209  LTSY_R                       0.054 
    M012P1_-  VOLM20_R                     0.057  VOLM16_R                     0.194 
    M012P1_-  VOLM12_R                     0.057  GP---_0R                         1 
    M012P1_-  R012_TP6                         1 
    M012P1_+  GS---_4R                      0.05  GS---_6R                      0.02 
    M012P1_+  LTSY_R                     0.05022  VOLM16_R                     0.194 
    M012P1_+  VOLM12_R                     0.057  GP---_0R                         1 
    M012P1_+  R012_TP6                         1  VOLM17_R                     0.201 
    M012P1_+  VOLM_8_R                    0.1209  GS---_5R                   0.03333 
    M012P1_+  INVEN_R                      0.203  AVEINV_R                   0.24833 
    M012P1_+  VOLM_7_R                     0.287  Obj                        0.13698 
    M012P2_1  VOLM_7_R                     0.096  Obj                         1.1999 
    M012P2_1  VOLM_1_R                    0.2814  AVEINV_R                    0.2388 
    M012P2_1  VOLM_9_R                    0.0686  VOLM_5_R                      0.05 
    M012P2_1  GS---_2R                      0.05  GS---_5R                   0.06667 
    M012P2_1  INVEN_R                    0.25415  VOLM18_R                    0.0686 
    M012P2_1  VOLM10_R                    0.1885  VOLM14_R                      0.05 
    M012P2_1  VOLM12_R                    0.1173  VOLM16_R                     0.096 
    M012P2_1  GP---_0R                         1  VOLM19_R                    0.1885 
    M012P2_1  LTSY_R                     0.05782  GS---_6R                      0.03 
    M012P2_1  GS---_1R                       0.1  VOLM_2_R                    0.1269 
    M012P2_1  R012_TP1                         1 
    M012P2_2  AVEINV_R                   0.25374  VOLM_2_R                    0.1269 
    M012P2_2  VOLM11_R                    0.1955  VOLM_1_R                    0.2814 
    M012P2_2  Obj                         1.1857  GS---_1R                       0.1 
    M012P2_2  R012_TP1                         1  GS---_6R                      0.03 
    M012P2_2  VOLM13_R                    0.1173  VOLM_5_R                      0.05 
    M012P2_2  VOLM15_R                      0.05  GS---_2R                      0.05 
    M012P2_2  INVEN_R                      0.377  GS---_5R                   0.03333 
    M012P2_2  LTSY_R                     0.05342  VOLM17_R                     0.096 
    M012P2_2  VOLM10_R                    0.0754  VOLM_7_R                     0.096 
    M012P2_2  VOLM20_R                    0.0754  GP---_0R                         1 
    M012P2_3  R012_TP2                         1  VOLM20_R                    0.1885 
    M012P2_3  GP---_0R                         1  Obj                        0.88399 
    M012P2_3  VOLM15_R                      0.05  VOLM10_R                    0.0686 
    M012P2_3  VOLM13_R                    0.1173  GS---_6R                      0.04 
    M012P2_3  LTSY_R                     0.05782  VOLM17_R                     0.096 
    M012P2_3  VOLM11_R                    0.1885  AVEINV_R                    0.2388 
    M012P2_3  VOLM_2_R                    0.2961  GS---_2R                       0.1 
    M012P2_3  VOLM_8_R                     0.096  GS---_5R                   0.03333 
    M012P2_3  VOLM19_R                    0.0686  VOLM_6_R                      0.05 
    M012P2_3  INVEN_R                     0.3393  VOLM_3_R                     0.132 
    M012P2_4  VOLM_8_R                     0.096  GS---_2R                       0.1 
    M012P2_4  VOLM_2_R                    0.2961  VOLM11_R                    0.0754 
    M012P2_4  AVEINV_R                   0.25374  VOLM18_R                     0.096 
    M012P2_4  GP---_0R                         1  LTSY_R                     0.05342 
    M012P2_4  Obj                        0.87414  R012_TP2                         1 
    M012P2_4  VOLM12_R                    0.1955  VOLM_6_R                      0.05 
    M012P2_4  VOLM16_R                      0.05  VOLM14_R                    0.1173 
    M012P2_4  VOLM_3_R                     0.132  INVEN_R                      0.343 
    M012P2_4  GS---_6R                      0.03 
    M012P2_5  VOLM_7_R                      0.05  INVEN_R                      0.343 
    M012P2_5  VOLM20_R                    0.0686  VOLM18_R                     0.096 
    M012P2_5  VOLM_9_R                     0.096  R012_TP3                         1 
    M012P2_5  VOLM11_R                    0.0686  GP---_0R                         1 
    M012P2_5  AVEINV_R                    0.2388  GS---_2R                      0.05 
    M012P2_5  LTSY_R                     0.05782  GS---_3R                      0.05 
    M012P2_5  VOLM12_R                    0.1885  GS---_6R                      0.04 
    M012P2_5  VOLM_3_R                     0.308  VOLM16_R                      0.05 
    M012P2_5  VOLM14_R                    0.1173  VOLM_4_R                    0.1311 
    M012P2_5  Obj                        0.64623 
    M012P2_6  Obj                         0.6401  VOLM15_R                    0.1173 
    M012P2_6  GP---_0R                         1  VOLM_3_R                     0.308 
    M012P2_6  VOLM12_R                    0.0754  VOLM13_R                    0.1955 
    M012P2_6  AVEINV_R                   0.25374  VOLM_4_R                    0.1311 
    M012P2_6  GS---_2R                      0.05  VOLM_9_R                     0.096 
    M012P2_6  INVEN_R                      0.339  VOLM17_R                      0.05 
    M012P2_6  GS---_3R                      0.05  VOLM_7_R                      0.05 
    M012P2_6  GS---_6R                      0.03  R012_TP3                         1 
    M012P2_6  VOLM19_R                     0.096  LTSY_R                     0.05342 
    M012P2_7  LTSY_R                     0.05782  R012_TP4                         1 
    M012P2_7  VOLM17_R                      0.05  VOLM13_R                    0.1885 
    M012P2_7  GS---_3R                       0.1  INVEN_R                      0.339 
    M012P2_7  VOLM_4_R                    0.3059  AVEINV_R                    0.2388 
    M012P2_7  VOLM10_R                     0.096  VOLM_5_R                    0.1287 
    M012P2_7  VOLM15_R                    0.1173  GS---_6R                      0.03 
    M012P2_7  VOLM19_R                     0.096  GP---_0R                         1 
    M012P2_7  VOLM12_R                    0.0686  Obj                        0.46975 
    M012P2_7  VOLM_8_R                      0.05 
    M012P2_8  VOLM_8_R                      0.05  GS---_6R                      0.03 
    M012P2_8  GP---_0R                         1  AVEINV_R                   0.25374 
    M012P2_8  VOLM16_R                    0.1173  VOLM13_R                    0.0754 
    M012P2_8  VOLM_4_R                    0.3059  GS---_3R                       0.1 
    M012P2_8  INVEN_R                      0.317  VOLM14_R                    0.1955 
    M012P2_8  R012_TP4                         1  VOLM10_R                     0.096 
    M012P2_8  VOLM18_R                      0.05  VOLM_5_R                    0.1287 
    M012P2_8  Obj                        0.46567  VOLM20_R                     0.096 
    M012P2_8  LTSY_R                     0.05342 
    M012P2_9  VOLM_6_R                    0.1251  VOLM_9_R                      0.05 
    M012P2_9  VOLM18_R                      0.05  GS---_6R                      0.03 
    M012P2_9  AVEINV_R                    0.2388  VOLM13_R                    0.0686 
    M012P2_9  GP---_0R                         1  VOLM_5_R                    0.3003 
    M012P2_9  VOLM11_R                     0.096  INVEN_R                      0.317 
    M012P2_9  VOLM16_R                    0.1173  LTSY_R                     0.05782 
    M012P2_9  Obj                        0.31886  GS---_4R                      0.05 
    M012P2_9  R012_TP5                         1  GS---_3R                      0.05 
    M012P2_9  VOLM14_R                    0.1885  VOLM20_R                     0.096 
    M012P2_A  GS---_3R                      0.05  Obj                        0.31592 
    M012P2_A  LTSY_R                     0.05342  R012_TP5                         1 
    M012P2_A  GS---_4R                      0.05  INVEN_R                      0.244 
    M012P2_A  VOLM11_R                     0.096  GP---_0R                         1 
    M012P2_A  GS---_6R                      0.03  AVEINV_R                   0.25374 
    M012P2_A  VOLM17_R                    0.1173  VOLM_5_R                    0.3003 
    M012P2_A  VOLM15_R                    0.1955  VOLM_6_R                    0.1251 
    M012P2_A  VOLM19_R                      0.05  VOLM_9_R                      0.05 
    M012P2_A  VOLM14_R                    0.0754 
    M012P2_B  VOLM15_R                    0.1885  VOLM_7_R                     0.123 
    M012P2_B  Obj                        0.20927  GS---_6R                      0.03 
    M012P2_B  AVEINV_R                    0.2388  INVEN_R                      0.244 
    M012P2_B  VOLM10_R                      0.05  LTSY_R                     0.05782 
    M012P2_B  VOLM_6_R                    0.2919  VOLM14_R                    0.0686 
    M012P2_B  GS---_4R                       0.1  VOLM19_R                      0.05 
    M012P2_B  VOLM17_R                    0.1173  GP---_0R                         1 
    M012P2_B  VOLM12_R                     0.096  R012_TP6                         1 
    M012P2_C  VOLM12_R                     0.096  VOLM20_R                      0.05 
    M012P2_C  R012_TP6                         1  Obj                        0.20784 
    M012P2_C  VOLM10_R                      0.05  VOLM_7_R                     0.123 
    M012P2_C  LTSY_R                     0.05342  VOLM16_R                    0.1955 
    M012P2_C  VOLM15_R                    0.0754  INVEN_R                      0.206 
    M012P2_C  VOLM_6_R                    0.2919  GP---_0R                         1 
    M012P2_C  GS---_6R                      0.03  GS---_4R                       0.1 
    M012P2_C  AVEINV_R                   0.25374  VOLM18_R                    0.1173 
    M012P2_D  VOLM18_R                    0.1173  AVEINV_R                    0.2388 
    M012P2_D  GS---_4R                      0.05  GS---_5R                   0.03333 
    M012P2_D  GP---_0R                         1  INVEN_R                      0.206 
    M012P2_D  VOLM11_R                      0.05  VOLM_8_R                    0.1209 
    M012P2_D  VOLM_7_R                     0.287  GS---_6R                      0.03 
    M012P2_D  VOLM16_R                    0.1885  VOLM15_R                    0.0686 
    M012P2_D  Obj                        0.14135  VOLM13_R                     0.096 
    M012P2_D  VOLM20_R                      0.05  R012_TP6                         1 
    M012P2_D  LTSY_R                     0.05782 
    M012P2_E  LTSY_R                     0.05342  VOLM16_R                    0.0754 
    M012P2_E  VOLM_7_R                     0.287  VOLM11_R                      0.05 
    M012P2_E  INVEN_R                      0.142  AVEINV_R                   0.25374 
    M012P2_E  GS---_5R                   0.03333  Obj                         0.1402 
    M012P2_E  VOLM19_R                    0.1173  VOLM17_R                    0.1955 
    M012P2_E  GS---_4R                      0.05  VOLM_8_R                    0.1209 
    M012P2_E  VOLM13_R                     0.096  GS---_6R                      0.03 
    M012P2_E  R012_TP6                         1  GP---_0R                         1 
    M037MN_1  R037_MN1                         1  Obj                     -0.0012632 
    M037RD_1  R037_RD1                         1  Obj                     -0.0010105 
    M037TF_1  R037_TM2                         1  INVEN_R                      0.142 
    M037TF_1  GS+++_3R                         1  VOLM_3_R                     0.389 
    M037TF_1  AVEINV_R                   0.18843  GP+++_0R                         1 
    M037TF_1  GS+++10R                         1  LTSY_R                     0.05157 
    M037TF_1  VOLM17_R                     0.361  Obj                        0.39119 
    M037TF_1  VOLM10_R                     0.361 
    M037TF_2  R037_TM2                         1  VOLM_3_R                     0.389 
    M037TF_2  VOLM11_R                     0.367  GS+++11R                         1 
    M037TF_2  Obj                        0.37769  GS+++_3R                         1 
    M037TF_2  GP+++_0R                         1  LTSY_R                     0.04587 
    M037TF_2  VOLM19_R                     0.367  AVEINV_R                   0.21075 
    M037TF_3  AVEINV_R                   0.18843  VOLM18_R                     0.361 
    M037TF_3  GP+++_0R                         1  GS+++_4R                         1 
    M037TF_3  Obj                        0.30569  VOLM11_R                     0.361 
    M037TF_3  R037_TM2                         1  VOLM_4_R                     0.402 
    M037TF_3  INVEN_R                      0.022  LTSY_R                     0.05157 
    M037TF_3  GS+++11R                         1 
    M037TF_4  AVEINV_R                   0.21075  VOLM_4_R                     0.402 
    M037TF_4  Obj                        0.29645  VOLM12_R                     0.367 
    M037TF_4  VOLM20_R                     0.367  LTSY_R                     0.04587 
    M037TF_4  INVEN_R                      0.367  R037_TM2                         1 
    M037TF_4  GS+++_4R                         1  GP+++_0R                         1 
    M037TF_4  GS+++12R                         1 
    M037TF_5  GS+++12R                         1  GS+++_5R                         1 
    M037TF_5  VOLM_5_R                     0.423  VOLM19_R                     0.361 
    M037TF_5  GP+++_0R                         1  LTSY_R                     0.05157 
    M037TF_5  AVEINV_R                   0.18843  VOLM12_R                     0.361 
    M037TF_5  R037_TM2                         1  Obj                        0.23234 
    M037TF_6  AVEINV_R                   0.21075  LTSY_R                     0.04587 
    M037TF_6  GS+++_5R                         1  INVEN_R                      0.361 
    M037TF_6  VOLM_5_R                     0.423  Obj                        0.22535 
    M037TF_6  GS+++13R                         1  R037_TM2                         1 
    M037TF_6  VOLM13_R                     0.367  GP+++_0R                         1 
    M037TF_7  VOLM20_R                     0.361  GS+++13R                         1 
    M037TF_7  R037_TM2                         1  GP+++_0R                         1 
    M037TF_7  Obj                        0.16521  AVEINV_R                   0.18843 
    M037TF_7  VOLM_6_R                      0.44  LTSY_R                     0.05157 
    M037TF_7  GS+++_6R                         1  VOLM13_R                     0.361 
    M037TF_7  INVEN_R                      0.361 
    M037TF_8  GP+++_0R                         1  VOLM14_R                     0.367 
    M037TF_8  INVEN_R                      0.324  GS+++14R                         1 
    M037TF_8  GS+++_6R                         1  Obj                        0.16042 
    M037TF_8  AVEINV_R                   0.21075  R037_TM2                         1 
    M037TF_8  VOLM_6_R                      0.44  LTSY_R                     0.04587 
    M037TF_9  GS+++_7R                         1  VOLM14_R                     0.361 
    M037TF_9  Obj                        0.12252  VOLM_7_R                     0.437 
    M037TF_9  GS+++14R                         1  R037_TM2                         1 
    M037TF_9  GP+++_0R                         1  AVEINV_R                   0.18843 
    M037TF_9  INVEN_R                      0.324  LTSY_R                     0.05157 
    M037TF_A  VOLM15_R                     0.367  VOLM_7_R                     0.437 
    M037TF_A  LTSY_R                     0.04587  INVEN_R                      0.267 
    M037TF_A  Obj                        0.11982  GS+++_7R                         1 
    M037TF_A  GS+++15R                         1  AVEINV_R                   0.21075 
    M037TF_A  GP+++_0R                         1  R037_TM2                         1 
    M037TF_B  AVEINV_R                   0.18843  GS+++_8R                         1 
    M037TF_B  Obj                       0.079982  R037_TM2                         1 
    M037TF_B  INVEN_R                      0.267  VOLM15_R                     0.361 
    M037TF_B  GP+++_0R                         1  GS+++15R                         1 
    M037TF_B  VOLM_8_R                     0.429  LTSY_R                     0.05157 
    M037TF_C  GS+++_8R                         1  GP+++_0R                         1 
    M037TF_C  VOLM16_R                     0.367  INVEN_R                      0.203 
    M037TF_C  R037_TM2                         1  VOLM_8_R                     0.429 
    M037TF_C  AVEINV_R                   0.21075  LTSY_R                     0.04587 
    M037TF_C  Obj                       0.078331 
    M037T1_1  VOLM14_R                     0.053  VOLM_3_R                     0.389 
    M037T1_1  INVEN_R                      0.142  R037_TM2                         1 
    M037T1_1  VOLM_7_R                     0.053  GS+++_3R                         1 
    M037T1_1  GP+++_0R                         1  AVEINV_R                   0.18143 
    M037T1_1  VOLM17_R                     0.339  LTSY_R                       0.056 
    M037T1_1  GS+++10R                         1  Obj                         0.4047 
    M037T1_1  VOLM10_R                     0.339 
    M037T1_2  VOLM_3_R                     0.389  GS+++_3R                         1 
    M037T1_2  AVEINV_R                   0.20187  VOLM11_R                     0.345 
    M037T1_2  LTSY_R                     0.04975  R037_TM2                         1 
    M037T1_2  VOLM_7_R                     0.053  Obj                        0.39242 
    M037T1_2  VOLM19_R                     0.345  GP+++_0R                         1 
    M037T1_2  GS+++11R                         1  VOLM15_R                     0.053 
    M037T1_3  Obj                        0.38433  VOLM_7_R                     0.053 
    M037T1_3  GS+++_3R                         1  R037_TM2                         1 
    M037T1_3  VOLM12_R                     0.377  GS+++12R                         1 
    M037T1_3  VOLM_3_R                     0.389  INVEN_R                      0.345 
    M037T1_3  AVEINV_R                   0.22133  GP+++_0R                         1 
    M037T1_3  VOLM16_R                     0.053  LTSY_R                     0.04778 
    M037T1_4  VOLM15_R                     0.053  GP+++_0R                         1 
    M037T1_4  VOLM11_R                     0.339  VOLM18_R                     0.339 
    M037T1_4  INVEN_R                      0.022  Obj                        0.31484 
    M037T1_4  GS+++_4R                         1  LTSY_R                       0.056 
    M037T1_4  R037_TM2                         1  VOLM_8_R                     0.053 
    M037T1_4  AVEINV_R                   0.18143  GS+++11R                         1 
    M037T1_4  VOLM_4_R                     0.402 
    M037T1_5  Obj                        0.30641  GS+++12R                         1 
    M037T1_5  AVEINV_R                   0.20187  VOLM16_R                     0.053 
    M037T1_5  R037_TM2                         1  INVEN_R                      0.345 
    M037T1_5  VOLM12_R                     0.345  VOLM_4_R                     0.402 
    M037T1_5  LTSY_R                     0.04975  VOLM_8_R                     0.053 
    M037T1_5  VOLM20_R                     0.345  GP+++_0R                         1 
    M037T1_5  GS+++_4R                         1 
    M037T1_6  INVEN_R                      0.339  VOLM_4_R                     0.402 
    M037T1_6  VOLM17_R                     0.053  VOLM_8_R                     0.053 
    M037T1_6  VOLM13_R                     0.377  GP+++_0R                         1 
    M037T1_6  R037_TM2                         1  AVEINV_R                   0.22133 
    M037T1_6  Obj                        0.30096  GS+++13R                         1 
    M037T1_6  GS+++_4R                         1  LTSY_R                     0.04778 
    M037T1_7  LTSY_R                       0.056  VOLM_9_R                     0.053 
    M037T1_7  Obj                        0.23852  GS+++12R                         1 
    M037T1_7  GS+++_5R                         1  VOLM19_R                     0.339 
    M037T1_7  VOLM12_R                     0.339  GP+++_0R                         1 
    M037T1_7  VOLM_5_R                     0.423  VOLM16_R                     0.053 
    M037T1_7  AVEINV_R                   0.18143  R037_TM2                         1 
    M037T1_8  GS+++13R                         1  VOLM13_R                     0.345 
    M037T1_8  VOLM17_R                     0.053  INVEN_R                      0.339 
    M037T1_8  GS+++_5R                         1  AVEINV_R                   0.20187 
    M037T1_8  R037_TM2                         1  VOLM_9_R                     0.053 
    M037T1_8  Obj                        0.23209  GP+++_0R                         1 
    M037T1_8  VOLM_5_R                     0.423  LTSY_R                     0.04975 
    M037T1_9  R037_TM2                         1  GP+++_0R                         1 
    M037T1_9  AVEINV_R                   0.22133  GS+++_5R                         1 
    M037T1_9  INVEN_R                      0.317  Obj                        0.22919 
    M037T1_9  GS+++14R                         1  LTSY_R                     0.04778 
    M037T1_9  VOLM14_R                     0.377  VOLM18_R                     0.053 
    M037T1_9  VOLM_9_R                     0.053  VOLM_5_R                     0.423 
    M037T1_A  LTSY_R                       0.056  VOLM13_R                     0.339 
    M037T1_A  VOLM_6_R                      0.44  INVEN_R                      0.339 
    M037T1_A  AVEINV_R                   0.18143  Obj                         0.1694 
    M037T1_A  GP+++_0R                         1  GS+++_6R                         1 
    M037T1_A  VOLM10_R                     0.053  VOLM20_R                     0.339 
    M037T1_A  R037_TM2                         1  GS+++13R                         1 
    M037T1_A  VOLM17_R                     0.053 
    M037T1_B  GS+++_6R                         1  VOLM14_R                     0.345 
    M037T1_B  VOLM10_R                     0.053  VOLM_6_R                      0.44 
    M037T1_B  INVEN_R                      0.317  GS+++14R                         1 
    M037T1_B  LTSY_R                     0.04975  Obj                        0.16499 
    M037T1_B  AVEINV_R                   0.20187  VOLM18_R                     0.053 
    M037T1_B  R037_TM2                         1  GP+++_0R                         1 
    M037T1_C  GS+++15R                         1  GP+++_0R                         1 
    M037T1_C  GS+++_6R                         1  VOLM_6_R                      0.44 
    M037T1_C  Obj                        0.16283  VOLM15_R                     0.377 
    M037T1_C  R037_TM2                         1  VOLM19_R                     0.053 
    M037T1_C  INVEN_R                      0.244  LTSY_R                     0.04778 
    M037T1_C  AVEINV_R                   0.22133  VOLM10_R                     0.053 
    M037T1_D  GP+++_0R                         1  GS+++_7R                         1 
    M037T1_D  LTSY_R                       0.056  VOLM14_R                     0.339 
    M037T1_D  R037_TM2                         1  GS+++14R                         1 
    M037T1_D  Obj                        0.12539  AVEINV_R                   0.18143 
    M037T1_D  VOLM_7_R                     0.437  INVEN_R                      0.317 
    M037T1_D  VOLM18_R                     0.053  VOLM11_R                     0.053 
    M037T1_E  VOLM19_R                     0.053  VOLM11_R                     0.053 
    M037T1_E  Obj                        0.12291  R037_TM2                         1 
    M037T1_E  GS+++15R                         1  GP+++_0R                         1 
    M037T1_E  GS+++_7R                         1  AVEINV_R                   0.20187 
    M037T1_E  VOLM15_R                     0.345  VOLM_7_R                     0.437 
    M037T1_E  LTSY_R                     0.04975  INVEN_R                      0.244 
    M037T1_F  INVEN_R                      0.206  VOLM20_R                     0.053 
    M037T1_F  LTSY_R                     0.04778  AVEINV_R                   0.22133 
    M037T1_F  VOLM_7_R                     0.437  VOLM11_R                     0.053 
    M037T1_F  GP+++_0R                         1  GS+++_7R                         1 
    M037T1_F  R037_TM2                         1  VOLM16_R                     0.377 
    M037T1_F  Obj                        0.12165 
    M037T1_G  GS+++_8R                         1  GS+++15R                         1 
    M037T1_G  INVEN_R                      0.244  VOLM19_R                     0.053 
    M037T1_G  Obj                       0.081923  VOLM15_R                     0.339 
    M037T1_G  GP+++_0R                         1  VOLM_8_R                     0.429 
    M037T1_G  R037_TM2                         1  LTSY_R                       0.056 
    M037T1_G  VOLM12_R                     0.053  AVEINV_R                   0.18143 
    M037T1_H  VOLM16_R                     0.345  VOLM20_R                     0.053 
    M037T1_H  Obj                       0.080439  GS+++_8R                         1 
    M037T1_H  VOLM_8_R                     0.429  R037_TM2                         1 
    M037T1_H  INVEN_R                      0.206  GP+++_0R                         1 
    M037T1_H  VOLM12_R                     0.053  AVEINV_R                   0.20187 
    M037T1_H  LTSY_R                     0.04975 
    M037T1_I  VOLM_8_R                     0.429  Obj                       0.079581 
    M037T1_I  R037_TM2                         1  GP+++_0R                         1 
    M037T1_I  INVEN_R                      0.142  AVEINV_R                   0.22133 
    M037T1_I  VOLM17_R                     0.377  GS+++_8R                         1 
    M037T1_I  VOLM12_R                     0.053  LTSY_R                     0.04778 
    M037T1_J  LTSY_R                       0.059  VOLM15_R                     0.057 
    M037T1_J  GS+++_3R                         1  AVEINV_R                   0.18343 
    M037T1_J  VOLM_3_R                     0.389  VOLM17_R                     0.356 
    M037T1_J  GP+++_0R                         1  GS+++10R                         1 
    M037T1_J  INVEN_R                      0.142  Obj                        0.40317 
    M037T1_J  R037_TM2                         1  VOLM_8_R                     0.057 
    M037T1_J  VOLM10_R                     0.356 
    M037T1_K  GS+++_3R                         1  GS+++11R                         1 
    M037T1_K  R037_TM2                         1  VOLM19_R                     0.362 
    M037T1_K  GP+++_0R                         1  VOLM_8_R                     0.057 
    M037T1_K  VOLM11_R                     0.362  LTSY_R                     0.05237 
    M037T1_K  Obj                        0.39026  VOLM16_R                     0.057 
    M037T1_K  AVEINV_R                   0.20575  VOLM_3_R                     0.389 
    M037T1_L  VOLM_3_R                     0.389  VOLM17_R                     0.057 
    M037T1_L  LTSY_R                     0.04944  GP+++_0R                         1 
    M037T1_L  INVEN_R                      0.362  VOLM_8_R                     0.057 
    M037T1_L  GS+++_3R                         1  GS+++12R                         1 
    M037T1_L  Obj                         0.3813  R037_TM2                         1 
    M037T1_L  VOLM12_R                     0.388  AVEINV_R                     0.226 
    M037T1_M  GS+++_4R                         1  LTSY_R                       0.059 
    M037T1_M  VOLM16_R                     0.057  VOLM_9_R                     0.057 
    M037T1_M  VOLM_4_R                     0.402  VOLM11_R                     0.356 
    M037T1_M  R037_TM2                         1  Obj                        0.31382 
    M037T1_M  GS+++11R                         1  GP+++_0R                         1 
    M037T1_M  INVEN_R                      0.022  AVEINV_R                   0.18343 
    M037T1_M  VOLM18_R                     0.356 
    M037T1_N  VOLM_4_R                     0.402  AVEINV_R                   0.20575 
    M037T1_N  INVEN_R                      0.362  GP+++_0R                         1 
    M037T1_N  VOLM20_R                     0.362  VOLM17_R                     0.057 
    M037T1_N  LTSY_R                     0.05237  Obj                        0.30495 
    M037T1_N  GS+++_4R                         1  R037_TM2                         1 
    M037T1_N  GS+++12R                         1  VOLM_9_R                     0.057 
    M037T1_N  VOLM12_R                     0.362 
    M037T1_O  GS+++13R                         1  R037_TM2                         1 
    M037T1_O  AVEINV_R                     0.226  INVEN_R                      0.356 
    M037T1_O  VOLM18_R                     0.057  VOLM_9_R                     0.057 
    M037T1_O  VOLM13_R                     0.388  GS+++_4R                         1 
    M037T1_O  VOLM_4_R                     0.402  Obj                        0.29892 
    M037T1_O  GP+++_0R                         1  LTSY_R                     0.04944 
    M037T1_P  VOLM_5_R                     0.423  VOLM10_R                     0.057 
    M037T1_P  VOLM12_R                     0.356  GS+++_5R                         1 
    M037T1_P  Obj                        0.23784  GP+++_0R                         1 
    M037T1_P  AVEINV_R                   0.18343  VOLM17_R                     0.057 
    M037T1_P  LTSY_R                       0.059  R037_TM2                         1 
    M037T1_P  VOLM19_R                     0.356  GS+++12R                         1 
    M037T1_Q  GS+++_5R                         1  VOLM_5_R                     0.423 
    M037T1_Q  VOLM13_R                     0.362  R037_TM2                         1 
    M037T1_Q  AVEINV_R                   0.20575  VOLM18_R                     0.057 
    M037T1_Q  INVEN_R                      0.356  LTSY_R                     0.05237 
    M037T1_Q  GP+++_0R                         1  VOLM10_R                     0.057 
    M037T1_Q  GS+++13R                         1  Obj                        0.23107 
    M037T1_R  VOLM14_R                     0.388  Obj                        0.22781 
    M037T1_R  LTSY_R                     0.04944  INVEN_R                      0.319 
    M037T1_R  GP+++_0R                         1  VOLM_5_R                     0.423 
    M037T1_R  R037_TM2                         1  VOLM19_R                     0.057 
    M037T1_R  GS+++_5R                         1  GS+++14R                         1 
    M037T1_R  AVEINV_R                     0.226  VOLM10_R                     0.057 
    M037T1_S  Obj                        0.16895  R037_TM2                         1 
    M037T1_S  VOLM13_R                     0.356  VOLM11_R                     0.057 
    M037T1_S  VOLM18_R                     0.057  GS+++13R                         1 
    M037T1_S  AVEINV_R                   0.18343  LTSY_R                       0.059 
    M037T1_S  GS+++_6R                         1  VOLM20_R                     0.356 
    M037T1_S  VOLM_6_R                      0.44  INVEN_R                      0.356 
    M037T1_S  GP+++_0R                         1 
    M037T1_T  GP+++_0R                         1  Obj                        0.16431 
    M037T1_T  VOLM11_R                     0.057  INVEN_R                      0.319 
    M037T1_T  VOLM19_R                     0.057  VOLM_6_R                      0.44 
    M037T1_T  LTSY_R                     0.05237  GS+++_6R                         1 
    M037T1_T  R037_TM2                         1  GS+++14R                         1 
    M037T1_T  AVEINV_R                   0.20575  VOLM14_R                     0.362 
    M037T1_U  Obj                         0.1619  GS+++_6R                         1 
    M037T1_U  VOLM20_R                     0.057  AVEINV_R                     0.226 
    M037T1_U  INVEN_R                      0.242  GS+++15R                         1 
    M037T1_U  VOLM11_R                     0.057  R037_TM2                         1 
    M037T1_U  VOLM_6_R                      0.44  VOLM15_R                     0.388 
    M037T1_U  LTSY_R                     0.04944  GP+++_0R                         1 
    M037T1_V  VOLM14_R                     0.356  INVEN_R                      0.319 
    M037T1_V  GS+++_7R                         1  AVEINV_R                   0.18343 
    M037T1_V  R037_TM2                         1  VOLM12_R                     0.057 
    M037T1_V  GS+++14R                         1  LTSY_R                       0.059 
    M037T1_V  GP+++_0R                         1  VOLM19_R                     0.057 
    M037T1_V  VOLM_7_R                     0.437  Obj                        0.12505 
    M037T1_W  AVEINV_R                   0.20575  R037_TM2                         1 
    M037T1_W  VOLM_7_R                     0.437  INVEN_R                      0.242 
    M037T1_W  GS+++_7R                         1  GP+++_0R                         1 
    M037T1_W  VOLM12_R                     0.057  LTSY_R                     0.05237 
    M037T1_W  GS+++15R                         1  Obj                        0.12245 
    M037T1_W  VOLM20_R                     0.057  VOLM15_R                     0.362 
    M037T1_X  GP+++_0R                         1  VOLM_7_R                     0.437 
    M037T1_X  Obj                        0.12094  AVEINV_R                     0.226 
    M037T1_X  INVEN_R                      0.203  R037_TM2                         1 
    M037T1_X  GS+++_7R                         1  VOLM16_R                     0.388 
    M037T1_X  LTSY_R                     0.04944  VOLM12_R                     0.057 
    M037T1_Y  Obj                       0.081696  GS+++_8R                         1 
    M037T1_Y  GP+++_0R                         1  VOLM20_R                     0.057 
    M037T1_Y  R037_TM2                         1  VOLM_8_R                     0.429 
    M037T1_Y  VOLM13_R                     0.057  LTSY_R                       0.059 
    M037T1_Y  INVEN_R                      0.242  AVEINV_R                   0.18343 
    M037T1_Y  GS+++15R                         1  VOLM15_R                     0.356 
    M037T1_Z  GP+++_0R                         1  AVEINV_R                   0.20575 
    M037T1_Z  VOLM16_R                     0.362  INVEN_R                      0.203 
    M037T1_Z  LTSY_R                     0.05237  GS+++_8R                         1 
    M037T1_Z  R037_TM2                         1  Obj                       0.080044 
    M037T1_Z  VOLM13_R                     0.057  VOLM_8_R                     0.429 
    M037T1_[  VOLM17_R                     0.388  AVEINV_R                     0.226 
    M037T1_[  R037_TM2                         1  Obj                       0.079178 
    M037T1_[  GP+++_0R                         1  INVEN_R                      0.142 
    M037T1_[  GS+++_8R                         1  VOLM13_R                     0.057 
    M037T1_[  LTSY_R                     0.04944  VOLM_8_R                     0.429 
    M037T2_1  GS+++11R                         1  R037_TM2                         1 
    M037T2_1  VOLM_3_R                     0.389  LTSY_R                     0.06112 
    M037T2_1  VOLM_7_R                      0.05  Obj                        0.40756 
    M037T2_1  GP+++_0R                         1  GS+++_3R                         1 
    M037T2_1  VOLM17_R                     0.096  VOLM11_R                     0.343 
    M037T2_1  VOLM_9_R                     0.096  VOLM19_R                     0.343 
    M037T2_1  AVEINV_R                   0.20162  VOLM15_R                      0.05 
    M037T2_2  AVEINV_R                   0.22111  VOLM_9_R                     0.096 
    M037T2_2  VOLM12_R                     0.377  VOLM18_R                     0.096 
    M037T2_2  INVEN_R                      0.343  GP+++_0R                         1 
    M037T2_2  GS+++_3R                         1  Obj                        0.39943 
    M037T2_2  VOLM_7_R                      0.05  VOLM_3_R                     0.389 
    M037T2_2  LTSY_R                     0.05811  GS+++12R                         1 
    M037T2_2  R037_TM2                         1  VOLM16_R                      0.05 
    M037T2_3  LTSY_R                     0.06112  VOLM16_R                      0.05 
    M037T2_3  R037_TM2                         1  GS+++12R                         1 
    M037T2_3  GS+++_4R                         1  Obj                        0.31665 
    M037T2_3  VOLM10_R                     0.096  VOLM_8_R                      0.05 
    M037T2_3  INVEN_R                      0.343  VOLM_4_R                     0.402 
    M037T2_3  VOLM20_R                     0.343  VOLM12_R                     0.343 
    M037T2_3  AVEINV_R                   0.20162  GP+++_0R                         1 
    M037T2_3  VOLM18_R                     0.096 
    M037T2_4  GP+++_0R                         1  AVEINV_R                   0.22111 
    M037T2_4  VOLM_4_R                     0.402  INVEN_R                      0.339 
    M037T2_4  VOLM10_R                     0.096  Obj                         0.3112 
    M037T2_4  VOLM_8_R                      0.05  GS+++_4R                         1 
    M037T2_4  R037_TM2                         1  GS+++13R                         1 
    M037T2_4  LTSY_R                     0.05811  VOLM19_R                     0.096 
    M037T2_4  VOLM17_R                      0.05  VOLM13_R                     0.377 
    M037T2_5  LTSY_R                     0.06112  GS+++13R                         1 
    M037T2_5  GS+++_5R                         1  INVEN_R                      0.339 
    M037T2_5  VOLM13_R                     0.343  AVEINV_R                   0.20162 
    M037T2_5  GP+++_0R                         1  VOLM19_R                     0.096 
    M037T2_5  VOLM17_R                      0.05  VOLM11_R                     0.096 
    M037T2_5  VOLM_5_R                     0.423  Obj                        0.23905 
    M037T2_5  VOLM_9_R                      0.05  R037_TM2                         1 
    M037T2_6  R037_TM2                         1  Obj                        0.23603 
    M037T2_6  VOLM_9_R                      0.05  VOLM14_R                     0.377 
    M037T2_6  LTSY_R                     0.05811  INVEN_R                      0.317 
    M037T2_6  VOLM20_R                     0.096  VOLM11_R                     0.096 
    M037T2_6  VOLM18_R                      0.05  GS+++_5R                         1 
    M037T2_6  GP+++_0R                         1  GS+++14R                         1 
    M037T2_6  VOLM_5_R                     0.423  AVEINV_R                   0.22111 
    M037T2_7  GP+++_0R                         1  GS+++14R                         1 
    M037T2_7  VOLM18_R                      0.05  VOLM20_R                     0.096 
    M037T2_7  R037_TM2                         1  VOLM_6_R                      0.44 
    M037T2_7  VOLM10_R                      0.05  INVEN_R                      0.317 
    M037T2_7  Obj                        0.16962  VOLM12_R                     0.096 
    M037T2_7  VOLM14_R                     0.343  LTSY_R                     0.06112 
    M037T2_7  AVEINV_R                   0.20162  GS+++_6R                         1 
    M037T2_8  GP+++_0R                         1  VOLM12_R                     0.096 
    M037T2_8  R037_TM2                         1  VOLM15_R                     0.377 
    M037T2_8  INVEN_R                      0.244  VOLM_6_R                      0.44 
    M037T2_8  GS+++_6R                         1  GS+++15R                         1 
    M037T2_8  VOLM19_R                      0.05  AVEINV_R                   0.22111 
    M037T2_8  VOLM10_R                      0.05  Obj                        0.16736 
    M037T2_8  LTSY_R                     0.05811 
    M037T2_9  VOLM11_R                      0.05  LTSY_R                     0.06112 
    M037T2_9  Obj                        0.12594  GS+++15R                         1 
    M037T2_9  AVEINV_R                   0.20162  GS+++_7R                         1 
    M037T2_9  INVEN_R                      0.244  VOLM15_R                     0.343 
    M037T2_9  VOLM13_R                     0.096  VOLM_7_R                     0.437 
    M037T2_9  R037_TM2                         1  GP+++_0R                         1 
    M037T2_9  VOLM19_R                      0.05 
    M037T2_A  VOLM11_R                      0.05  AVEINV_R                   0.22111 
    M037T2_A  VOLM20_R                      0.05  INVEN_R                      0.206 
    M037T2_A  Obj                        0.12473  VOLM16_R                     0.377 
    M037T2_A  R037_TM2                         1  LTSY_R                     0.05811 
    M037T2_A  GS+++_7R                         1  GP+++_0R                         1 
    M037T2_A  VOLM13_R                     0.096  VOLM_7_R                     0.437 
    M037T2_B  LTSY_R                     0.06112  INVEN_R                      0.206 
    M037T2_B  VOLM12_R                      0.05  VOLM20_R                      0.05 
    M037T2_B  VOLM16_R                     0.343  VOLM14_R                     0.096 
    M037T2_B  Obj                       0.082506  VOLM_8_R                     0.429 
    M037T2_B  AVEINV_R                   0.20162  R037_TM2                         1 
    M037T2_B  GS+++_8R                         1  GP+++_0R                         1 
    M037T2_C  GS+++_8R                         1  GP+++_0R                         1 
    M037T2_C  INVEN_R                      0.142  AVEINV_R                   0.22111 
    M037T2_C  VOLM17_R                     0.377  LTSY_R                     0.05811 
    M037T2_C  R037_TM2                         1  VOLM14_R                     0.096 
    M037T2_C  VOLM_8_R                     0.429  VOLM12_R                      0.05 
    M037T2_C  Obj                       0.081696 
    M037PF_1  VOLM_3_R                    0.2723  GP---_0R                         1 
    M037PF_1  LTSY_R                     0.05185  Obj                        0.53454 
    M037PF_1  VOLM17_R                   0.24548  VOLM_4_R                    0.1206 
    M037PF_1  GS---_6R                      0.03  GS---_5R                   0.03333 
    M037PF_1  VOLM10_R                   0.24548  VOLM18_R                   0.11744 
    M037PF_1  INVEN_R                      0.142  R037_TP2                         1 
    M037PF_1  GS---_3R                      0.05  AVEINV_R                   0.20521 
    M037PF_1  VOLM11_R                   0.11744  GS---_2R                      0.05 
    M037PF_2  VOLM11_R                   0.24956  GS---_3R                      0.05 
    M037PF_2  GP---_0R                         1  VOLM_3_R                    0.2723 
    M037PF_2  LTSY_R                     0.04659  Obj                        0.52207 
    M037PF_2  INVEN_R                     0.2541  GS---_6R                      0.04 
    M037PF_2  VOLM_4_R                    0.1206  GS---_2R                      0.05 
    M037PF_2  VOLM12_R                    0.1232  AVEINV_R                   0.22615 
    M037PF_2  VOLM20_R                    0.1232  R037_TP2                         1 
    M037PF_2  VOLM19_R                   0.24956 
    M037PF_3  VOLM19_R                   0.11744  VOLM18_R                   0.24548 
    M037PF_3  AVEINV_R                   0.20521  GP---_0R                         1 
    M037PF_3  VOLM12_R                   0.11744  R037_TP2                         1 
    M037PF_3  VOLM_5_R                    0.1269  INVEN_R                      0.022 
    M037PF_3  LTSY_R                     0.05185  VOLM11_R                   0.24548 
    M037PF_3  VOLM_4_R                    0.2814  Obj                        0.39773 
    M037PF_3  GS---_3R                       0.1  GS---_6R                      0.04 
    M037PF_4  GS---_3R                       0.1  LTSY_R                     0.04659 
    M037PF_4  VOLM_4_R                    0.2814  VOLM_5_R                    0.1269 
    M037PF_4  Obj                        0.38918  VOLM13_R                    0.1232 
    M037PF_4  GP---_0R                         1  R037_TP2                         1 
    M037PF_4  VOLM20_R                   0.24956  INVEN_R                      0.367 
    M037PF_4  AVEINV_R                   0.22615  VOLM12_R                   0.24956 
    M037PF_4  GS---_6R                      0.03 
    M037PF_5  INVEN_R                    0.24222  Obj                        0.28935 
    M037PF_5  R037_TP2                         1  AVEINV_R                   0.20521 
    M037PF_5  LTSY_R                     0.05185  GS---_4R                      0.05 
    M037PF_5  GS---_3R                      0.05  VOLM20_R                   0.11744 
    M037PF_5  GS---_6R                      0.04  VOLM_6_R                     0.132 
    M037PF_5  GP---_0R                         1  VOLM19_R                   0.24548 
    M037PF_5  VOLM_5_R                    0.2961  VOLM13_R                   0.11744 
    M037PF_5  VOLM12_R                   0.24548 
    M037PF_6  Obj                        0.28307  GS---_3R                      0.05 
    M037PF_6  GS---_6R                      0.02  VOLM13_R                   0.24956 
    M037PF_6  AVEINV_R                   0.22615  VOLM_6_R                     0.132 
    M037PF_6  VOLM_5_R                    0.2961  LTSY_R                     0.04659 
    M037PF_6  R037_TP2                         1  GS---_4R                      0.05 
    M037PF_6  GP---_0R                         1  INVEN_R                      0.361 
    M037PF_6  VOLM14_R                    0.1232 
    M037PF_7  INVEN_R                      0.361  LTSY_R                     0.05185 
    M037PF_7  GP---_0R                         1  GS---_4R                       0.1 
    M037PF_7  GS---_6R                      0.03  R037_TP2                         1 
    M037PF_7  VOLM_6_R                     0.308  AVEINV_R                   0.20521 
    M037PF_7  VOLM14_R                   0.11744  VOLM_7_R                    0.1311 
    M037PF_7  Obj                        0.20463  VOLM20_R                   0.24548 
    M037PF_7  VOLM13_R                   0.24548 
    M037PF_8  Obj                        0.20061  VOLM14_R                   0.24956 
    M037PF_8  GP---_0R                         1  VOLM_6_R                     0.308 
    M037PF_8  R037_TP2                         1  GS---_4R                       0.1 
    M037PF_8  GS---_6R                      0.02  AVEINV_R                   0.22615 
    M037PF_8  VOLM_7_R                    0.1311  VOLM15_R                    0.1232 
    M037PF_8  LTSY_R                     0.04659  INVEN_R                      0.324 
    M037PF_9  Obj                        0.14528  VOLM14_R                   0.24548 
    M037PF_9  GP---_0R                         1  INVEN_R                      0.324 
    M037PF_9  VOLM_8_R                    0.1287  R037_TP2                         1 
    M037PF_9  VOLM15_R                   0.11744  GS---_4R                      0.05 
    M037PF_9  LTSY_R                     0.05185  AVEINV_R                   0.20521 
    M037PF_9  VOLM_7_R                    0.3059  GS---_5R                   0.03333 
    M037PF_9  GS---_6R                      0.02 
    M037PF_A  AVEINV_R                   0.22615  GS---_4R                      0.05 
    M037PF_A  R037_TP2                         1  GS---_5R                   0.03333 
    M037PF_A  LTSY_R                     0.04659  VOLM_7_R                    0.3059 
    M037PF_A  VOLM16_R                    0.1232  GS---_6R                      0.02 
    M037PF_A  VOLM_8_R                    0.1287  INVEN_R                      0.267 
    M037PF_A  Obj                        0.14286  VOLM15_R                   0.24956 
    M037PF_A  GP---_0R                         1 
    M037PF_B  AVEINV_R                   0.20521  R037_TP2                         1 
    M037PF_B  Obj                       0.095506  GS---_5R                   0.06667 
    M037PF_B  VOLM16_R                   0.11744  VOLM_8_R                    0.3003 
    M037PF_B  GS---_6R                      0.02  INVEN_R                      0.267 
    M037PF_B  GP---_0R                         1  VOLM_9_R                    0.1251 
    M037PF_B  VOLM15_R                   0.24548  LTSY_R                     0.05185 
    M037PF_C  Obj                       0.094017  GS---_5R                   0.06667 
    M037PF_C  VOLM_9_R                    0.1251  INVEN_R                      0.203 
    M037PF_C  AVEINV_R                   0.22615  VOLM17_R                    0.1232 
    M037PF_C  LTSY_R                     0.04659  R037_TP2                         1 
    M037PF_C  GS---_6R                      0.02  VOLM_8_R                    0.3003 
    M037PF_C  VOLM16_R                   0.24956  GP---_0R                         1 
    M037P1_1  GP---_0R                         1  GS---_5R                   0.03333 
    M037P1_1  VOLM_7_R                     0.053  VOLM14_R                     0.053 
    M037P1_1  Obj                        0.54622  LTSY_R                     0.05643 
    M037P1_1  GS---_2R                      0.05  VOLM_4_R                    0.1206 
    M037P1_1  GS---_6R                      0.03  VOLM11_R                    0.1725 
    M037P1_1  VOLM18_R                    0.1725  GS---_3R                      0.05 
    M037P1_1  AVEINV_R                   0.20607  R037_TP2                         1 
    M037P1_1  VOLM10_R                    0.1695  INVEN_R                      0.142 
    M037P1_1  VOLM17_R                    0.1695  VOLM_3_R                    0.2723 
    M037P1_2  R037_TP2                         1  VOLM_7_R                     0.053 
    M037P1_2  VOLM_3_R                    0.2723  GP---_0R                         1 
    M037P1_2  INVEN_R                    0.28275  VOLM12_R                    0.1885 
    M037P1_2  GS---_3R                      0.05  AVEINV_R                   0.22544 
    M037P1_2  VOLM11_R                    0.1725  VOLM15_R                     0.053 
    M037P1_2  VOLM19_R                    0.1725  GS---_2R                      0.05 
    M037P1_2  VOLM_4_R                    0.1206  GS---_6R                      0.04 
    M037P1_2  Obj                        0.53569  LTSY_R                     0.05175 
    M037P1_2  VOLM20_R                    0.1885 
    M037P1_3  AVEINV_R                   0.24306  INVEN_R                      0.345 
    M037P1_3  VOLM13_R                    0.1955  GS---_2R                      0.05 
    M037P1_3  VOLM_4_R                    0.1206  VOLM_3_R                    0.2723 
    M037P1_3  LTSY_R                     0.04856  VOLM16_R                     0.053 
    M037P1_3  GS---_6R                      0.02  Obj                        0.52742 
    M037P1_3  VOLM_7_R                     0.053  GP---_0R                         1 
    M037P1_3  VOLM12_R                    0.1885  GS---_3R                      0.05 
    M037P1_3  R037_TP2                         1 
    M037P1_4  AVEINV_R                   0.20607  GS---_3R                       0.1 
    M037P1_4  VOLM15_R                     0.053  R037_TP2                         1 
    M037P1_4  VOLM12_R                    0.1725  VOLM_5_R                    0.1269 
    M037P1_4  VOLM_8_R                     0.053  VOLM19_R                    0.1725 
    M037P1_4  GS---_6R                      0.04  LTSY_R                     0.05643 
    M037P1_4  VOLM11_R                    0.1695  GP---_0R                         1 
    M037P1_4  INVEN_R                      0.022  VOLM_4_R                    0.2814 
    M037P1_4  VOLM18_R                    0.1695  Obj                        0.40563 
    M037P1_5  INVEN_R                      0.345  VOLM20_R                    0.1725 
    M037P1_5  GP---_0R                         1  VOLM_4_R                    0.2814 
    M037P1_5  VOLM12_R                    0.1725  VOLM13_R                    0.1885 
    M037P1_5  LTSY_R                     0.05175  VOLM_5_R                    0.1269 
    M037P1_5  Obj                         0.3982  VOLM16_R                     0.053 
    M037P1_5  VOLM_8_R                     0.053  R037_TP2                         1 
    M037P1_5  AVEINV_R                   0.22544  GS---_6R                      0.03 
    M037P1_5  GS---_3R                       0.1 
    M037P1_6  R037_TP2                         1  GS---_3R                       0.1 
    M037P1_6  VOLM_5_R                    0.1269  VOLM_8_R                     0.053 
    M037P1_6  VOLM13_R                    0.1885  AVEINV_R                   0.24306 
    M037P1_6  Obj                        0.39314  LTSY_R                     0.04856 
    M037P1_6  VOLM14_R                    0.1955  GS---_6R                      0.02 
    M037P1_6  VOLM_4_R                    0.2814  INVEN_R                      0.339 
    M037P1_6  VOLM17_R                     0.053  GP---_0R                         1 
    M037P1_7  VOLM_9_R                     0.053  INVEN_R                    0.25875 
    M037P1_7  VOLM12_R                    0.1695  R037_TP2                         1 
    M037P1_7  LTSY_R                     0.05643  VOLM19_R                    0.1695 
    M037P1_7  GS---_3R                      0.05  AVEINV_R                   0.20607 
    M037P1_7  VOLM13_R                    0.1725  VOLM_5_R                    0.2961 
    M037P1_7  GS---_4R                      0.05  VOLM16_R                     0.053 
    M037P1_7  GS---_6R                      0.04  Obj                        0.29469 
    M037P1_7  VOLM20_R                    0.1725  VOLM_6_R                     0.132 
    M037P1_7  GP---_0R                         1 
    M037P1_8  VOLM_6_R                     0.132  GS---_4R                      0.05 
    M037P1_8  VOLM13_R                    0.1725  GS---_6R                      0.02 
    M037P1_8  VOLM_5_R                    0.2961  AVEINV_R                   0.22544 
    M037P1_8  VOLM17_R                     0.053  LTSY_R                     0.05175 
    M037P1_8  GS---_3R                      0.05  INVEN_R                      0.339 
    M037P1_8  Obj                        0.28932  GP---_0R                         1 
    M037P1_8  VOLM14_R                    0.1885  VOLM_9_R                     0.053 
    M037P1_8  R037_TP2                         1 
    M037P1_9  R037_TP2                         1  VOLM14_R                    0.1885 
    M037P1_9  GS---_4R                      0.05  VOLM_5_R                    0.2961 
    M037P1_9  INVEN_R                      0.317  VOLM_9_R                     0.053 
    M037P1_9  GS---_3R                      0.05  LTSY_R                     0.04856 
    M037P1_9  GS---_6R                      0.02  AVEINV_R                   0.24306 
    M037P1_9  VOLM_6_R                     0.132  Obj                        0.28629 
    M037P1_9  VOLM15_R                    0.1955  VOLM18_R                     0.053 
    M037P1_9  GP---_0R                         1 
    M037P1_A  GP---_0R                         1  VOLM10_R                     0.053 
    M037P1_A  VOLM20_R                    0.1695  Obj                         0.2081 
    M037P1_A  LTSY_R                     0.05643  AVEINV_R                   0.20607 
    M037P1_A  GS---_6R                      0.03  GS---_4R                       0.1 
    M037P1_A  INVEN_R                      0.339  R037_TP2                         1 
    M037P1_A  VOLM_6_R                     0.308  VOLM14_R                    0.1725 
    M037P1_A  VOLM13_R                    0.1695  VOLM17_R                     0.053 
    M037P1_A  VOLM_7_R                    0.1311 
    M037P1_B  VOLM_7_R                    0.1311  VOLM18_R                     0.053 
    M037P1_B  GS---_4R                       0.1  LTSY_R                     0.05175 
    M037P1_B  VOLM14_R                    0.1725  VOLM10_R                     0.053 
    M037P1_B  AVEINV_R                   0.22544  R037_TP2                         1 
    M037P1_B  GS---_6R                      0.02  VOLM15_R                    0.1885 
    M037P1_B  GP---_0R                         1  INVEN_R                      0.317 
    M037P1_B  Obj                        0.20485  VOLM_6_R                     0.308 
    M037P1_C  GS---_4R                       0.1  INVEN_R                      0.244 
    M037P1_C  VOLM16_R                    0.1955  VOLM15_R                    0.1885 
    M037P1_C  LTSY_R                     0.04856  VOLM_7_R                    0.1311 
    M037P1_C  GP---_0R                         1  VOLM10_R                     0.053 
    M037P1_C  AVEINV_R                   0.24306  VOLM_6_R                     0.308 
    M037P1_C  GS---_6R                      0.02  VOLM19_R                     0.053 
    M037P1_C  Obj                         0.2026  R037_TP2                         1 
    M037P1_D  LTSY_R                     0.05643  VOLM15_R                    0.1725 
    M037P1_D  VOLM18_R                     0.053  INVEN_R                      0.317 
    M037P1_D  VOLM_8_R                    0.1287  Obj                        0.14778 
    M037P1_D  VOLM_7_R                    0.3059  VOLM14_R                    0.1695 
    M037P1_D  GS---_6R                      0.02  AVEINV_R                   0.20607 
    M037P1_D  GS---_4R                      0.05  GP---_0R                         1 
    M037P1_D  R037_TP2                         1  VOLM11_R                     0.053 
    M037P1_D  GS---_5R                   0.03333 
    M037P1_E  VOLM15_R                    0.1725  GS---_5R                   0.03333 
    M037P1_E  VOLM16_R                    0.1885  VOLM11_R                     0.053 
    M037P1_E  VOLM19_R                     0.053  VOLM_7_R                    0.3059 
    M037P1_E  GS---_4R                      0.05  R037_TP2                         1 
    M037P1_E  Obj                        0.14573  INVEN_R                      0.244 
    M037P1_E  VOLM_8_R                    0.1287  GS---_6R                      0.02 
    M037P1_E  GP---_0R                         1  AVEINV_R                   0.22544 
    M037P1_E  LTSY_R                     0.05175 
    M037P1_F  INVEN_R                      0.206  VOLM_8_R                    0.1287 
    M037P1_F  GP---_0R                         1  GS---_6R                      0.02 
    M037P1_F  VOLM16_R                    0.1885  Obj                        0.14438 
    M037P1_F  R037_TP2                         1  VOLM17_R                    0.1955 
    M037P1_F  GS---_4R                      0.05  VOLM_7_R                    0.3059 
    M037P1_F  VOLM11_R                     0.053  GS---_5R                   0.03333 
    M037P1_F  LTSY_R                     0.04856  VOLM20_R                     0.053 
    M037P1_F  AVEINV_R                   0.24306 
    M037P1_G  VOLM15_R                    0.1695  GS---_6R                      0.02 
    M037P1_G  VOLM12_R                     0.053  LTSY_R                     0.05643 
    M037P1_G  GS---_5R                   0.06667  VOLM_8_R                    0.3003 
    M037P1_G  VOLM_9_R                    0.1251  INVEN_R                      0.244 
    M037P1_G  VOLM19_R                     0.053  R037_TP2                         1 
    M037P1_G  AVEINV_R                   0.20607  VOLM16_R                    0.1725 
    M037P1_G  GP---_0R                         1  Obj                       0.097203 
    M037P1_H  INVEN_R                      0.206  LTSY_R                     0.05175 
    M037P1_H  VOLM_8_R                    0.3003  GS---_5R                   0.06667 
    M037P1_H  GP---_0R                         1  VOLM_9_R                    0.1251 
    M037P1_H  R037_TP2                         1  VOLM16_R                    0.1725 
    M037P1_H  VOLM17_R                    0.1885  AVEINV_R                   0.22544 
    M037P1_H  VOLM12_R                     0.053  VOLM20_R                     0.053 
    M037P1_H  Obj                       0.095975  GS---_6R                      0.02 
    M037P1_I  VOLM18_R                    0.1955  VOLM12_R                     0.053 
    M037P1_I  Obj                       0.095433  GS---_6R                      0.02 
    M037P1_I  GS---_5R                   0.06667  LTSY_R                     0.04856 
    M037P1_I  R037_TP2                         1  VOLM_9_R                    0.1251 
    M037P1_I  GP---_0R                         1  AVEINV_R                   0.24306 
    M037P1_I  INVEN_R                      0.142  VOLM17_R                    0.1885 
    M037P1_I  VOLM_8_R                    0.3003 
    M037P1_J  VOLM17_R                     0.178  VOLM_4_R                    0.1206 
    M037P1_J  VOLM_8_R                     0.057  GS---_5R                   0.03333 
    M037P1_J  INVEN_R                      0.142  LTSY_R                     0.05943 
    M037P1_J  AVEINV_R                   0.20929  GP---_0R                         1 
    M037P1_J  GS---_6R                      0.03  VOLM10_R                     0.178 
    M037P1_J  R037_TP2                         1  VOLM18_R                     0.181 
    M037P1_J  Obj                        0.54437  VOLM11_R                     0.181 
    M037P1_J  VOLM_3_R                    0.2723  GS---_3R                      0.05 
    M037P1_J  GS---_2R                      0.05  VOLM15_R                     0.057 
    M037P1_K  Obj                        0.53309  VOLM20_R                     0.194 
    M037P1_K  VOLM_3_R                    0.2723  VOLM12_R                     0.194 
    M037P1_K  R037_TP2                         1  VOLM19_R                     0.181 
    M037P1_K  VOLM_4_R                    0.1206  GP---_0R                         1 
    M037P1_K  INVEN_R                      0.291  VOLM16_R                     0.057 
    M037P1_K  VOLM11_R                     0.181  GS---_2R                      0.05 
    M037P1_K  LTSY_R                       0.054  GS---_6R                      0.04 
    M037P1_K  AVEINV_R                      0.23  VOLM_8_R                     0.057 
    M037P1_K  GS---_3R                      0.05 
    M037P1_L  GS---_3R                      0.05  VOLM12_R                     0.194 
    M037P1_L  R037_TP2                         1  VOLM_3_R                    0.2723 
    M037P1_L  AVEINV_R                   0.24833  GP---_0R                         1 
    M037P1_L  VOLM_8_R                     0.057  GS---_2R                      0.05 
    M037P1_L  LTSY_R                     0.05022  VOLM13_R                     0.201 
    M037P1_L  VOLM_4_R                    0.1206  GS---_6R                      0.02 
    M037P1_L  VOLM17_R                     0.057  INVEN_R                      0.362 
    M037P1_L  Obj                        0.52428 
    M037P1_M  VOLM12_R                     0.181  Obj                        0.40438 
    M037P1_M  GS---_6R                      0.04  VOLM_4_R                    0.2814 
    M037P1_M  INVEN_R                      0.022  GS---_3R                       0.1 
    M037P1_M  AVEINV_R                   0.20929  VOLM16_R                     0.057 
    M037P1_M  VOLM_9_R                     0.057  VOLM_5_R                    0.1269 
    M037P1_M  GP---_0R                         1  VOLM19_R                     0.181 
    M037P1_M  LTSY_R                     0.05943  R037_TP2                         1 
    M037P1_M  VOLM18_R                     0.178  VOLM11_R                     0.178 
    M037P1_N  GS---_6R                      0.03  GP---_0R                         1 
    M037P1_N  VOLM_5_R                    0.1269  VOLM_4_R                    0.2814 
    M037P1_N  Obj                        0.39644  VOLM17_R                     0.057 
    M037P1_N  AVEINV_R                      0.23  VOLM12_R                     0.181 
    M037P1_N  VOLM13_R                     0.194  GS---_3R                       0.1 
    M037P1_N  INVEN_R                      0.362  LTSY_R                       0.054 
    M037P1_N  VOLM_9_R                     0.057  R037_TP2                         1 
    M037P1_N  VOLM20_R                     0.181 
    M037P1_O  VOLM_9_R                     0.057  VOLM18_R                     0.057 
    M037P1_O  R037_TP2                         1  VOLM13_R                     0.194 
    M037P1_O  GS---_6R                      0.02  VOLM14_R                     0.201 
    M037P1_O  GS---_3R                       0.1  VOLM_5_R                    0.1269 
    M037P1_O  LTSY_R                     0.05022  AVEINV_R                   0.24833 
    M037P1_O  VOLM_4_R                    0.2814  Obj                        0.39101 
    M037P1_O  GP---_0R                         1  INVEN_R                      0.356 
    M037P1_P  VOLM20_R                     0.181  GP---_0R                         1 
    M037P1_P  R037_TP2                         1  VOLM17_R                     0.057 
    M037P1_P  LTSY_R                     0.05943  VOLM13_R                     0.181 
    M037P1_P  VOLM12_R                     0.178  GS---_3R                      0.05 
    M037P1_P  GS---_4R                      0.05  AVEINV_R                   0.20929 
    M037P1_P  GS---_6R                      0.04  VOLM_6_R                     0.132 
    M037P1_P  Obj                        0.29385  INVEN_R                     0.2715 
    M037P1_P  VOLM19_R                     0.178  VOLM10_R                     0.057 
    M037P1_P  VOLM_5_R                    0.2961 
    M037P1_Q  VOLM14_R                     0.194  GP---_0R                         1 
    M037P1_Q  LTSY_R                       0.054  VOLM_6_R                     0.132 
    M037P1_Q  VOLM_5_R                    0.2961  VOLM10_R                     0.057 
    M037P1_Q  INVEN_R                      0.356  VOLM18_R                     0.057 
    M037P1_Q  VOLM13_R                     0.181  GS---_6R                      0.02 
    M037P1_Q  Obj                        0.28811  GS---_3R                      0.05 
    M037P1_Q  AVEINV_R                      0.23  GS---_4R                      0.05 
    M037P1_Q  R037_TP2                         1 
    M037P1_R  VOLM15_R                     0.201  GS---_4R                      0.05 
    M037P1_R  VOLM14_R                     0.194  GS---_3R                      0.05 
    M037P1_R  VOLM10_R                     0.057  Obj                        0.28486 
    M037P1_R  GS---_6R                      0.02  VOLM_5_R                    0.2961 
    M037P1_R  VOLM_6_R                     0.132  LTSY_R                     0.05022 
    M037P1_R  INVEN_R                      0.319  R037_TP2                         1 
    M037P1_R  VOLM19_R                     0.057  GP---_0R                         1 
    M037P1_R  AVEINV_R                   0.24833 
    M037P1_S  VOLM_7_R                    0.1311  AVEINV_R                   0.20929 
    M037P1_S  VOLM20_R                     0.178  GS---_4R                       0.1 
    M037P1_S  GP---_0R                         1  R037_TP2                         1 
    M037P1_S  VOLM18_R                     0.057  GS---_6R                      0.03 
    M037P1_S  INVEN_R                      0.356  VOLM14_R                     0.181 
    M037P1_S  LTSY_R                     0.05943  Obj                        0.20752 
    M037P1_S  VOLM_6_R                     0.308  VOLM11_R                     0.057 
    M037P1_S  VOLM13_R                     0.178 
    M037P1_T  Obj                        0.20404  VOLM_6_R                     0.308 
    M037P1_T  VOLM15_R                     0.194  LTSY_R                       0.054 
    M037P1_T  R037_TP2                         1  VOLM14_R                     0.181 
    M037P1_T  VOLM11_R                     0.057  GS---_6R                      0.02 
    M037P1_T  INVEN_R                      0.319  GP---_0R                         1 
    M037P1_T  VOLM19_R                     0.057  GS---_4R                       0.1 
    M037P1_T  VOLM_7_R                    0.1311  AVEINV_R                      0.23 
    M037P1_U  VOLM_7_R                    0.1311  Obj                        0.20164 
    M037P1_U  R037_TP2                         1  AVEINV_R                   0.24833 
    M037P1_U  INVEN_R                      0.242  LTSY_R                     0.05022 
    M037P1_U  VOLM16_R                     0.201  GP---_0R                         1 
    M037P1_U  GS---_4R                       0.1  VOLM20_R                     0.057 
    M037P1_U  GS---_6R                      0.02  VOLM15_R                     0.194 
    M037P1_U  VOLM_6_R                     0.308  VOLM11_R                     0.057 
    M037P1_V  VOLM_7_R                    0.3059  VOLM12_R                     0.057 
    M037P1_V  VOLM_8_R                    0.1287  GS---_6R                      0.02 
    M037P1_V  VOLM15_R                     0.181  INVEN_R                      0.319 
    M037P1_V  LTSY_R                     0.05943  VOLM14_R                     0.178 
    M037P1_V  AVEINV_R                   0.20929  R037_TP2                         1 
    M037P1_V  GP---_0R                         1  GS---_4R                      0.05 
    M037P1_V  VOLM19_R                     0.057  GS---_5R                   0.03333 
    M037P1_V  Obj                        0.14738 
    M037P1_W  Obj                        0.14518  VOLM_8_R                    0.1287 
    M037P1_W  R037_TP2                         1  VOLM12_R                     0.057 
    M037P1_W  GS---_5R                   0.03333  GS---_6R                      0.02 
    M037P1_W  VOLM16_R                     0.194  INVEN_R                      0.242 
    M037P1_W  VOLM20_R                     0.057  VOLM15_R                     0.181 
    M037P1_W  VOLM_7_R                    0.3059  LTSY_R                       0.054 
    M037P1_W  GS---_4R                      0.05  AVEINV_R                      0.23 
    M037P1_W  GP---_0R                         1 
    M037P1_X  GP---_0R                         1  GS---_4R                      0.05 
    M037P1_X  VOLM_7_R                    0.3059  VOLM12_R                     0.057 
    M037P1_X  GS---_6R                      0.02  Obj                        0.14364 
    M037P1_X  VOLM_8_R                    0.1287  R037_TP2                         1 
    M037P1_X  VOLM17_R                     0.201  GS---_5R                   0.03333 
    M037P1_X  VOLM16_R                     0.194  LTSY_R                     0.05022 
    M037P1_X  INVEN_R                      0.203  AVEINV_R                   0.24833 
    M037P1_Y  VOLM15_R                     0.178  VOLM_8_R                    0.3003 
    M037P1_Y  GP---_0R                         1  Obj                       0.096931 
    M037P1_Y  GS---_6R                      0.02  AVEINV_R                   0.20929 
    M037P1_Y  VOLM13_R                     0.057  VOLM20_R                     0.057 
    M037P1_Y  INVEN_R                      0.242  LTSY_R                     0.05943 
    M037P1_Y  R037_TP2                         1  VOLM16_R                     0.181 
    M037P1_Y  GS---_5R                   0.06667  VOLM_9_R                    0.1251 
    M037P1_Z  VOLM17_R                     0.194  INVEN_R                      0.203 
    M037P1_Z  LTSY_R                       0.054  GS---_5R                   0.06667 
    M037P1_Z  R037_TP2                         1  VOLM13_R                     0.057 
    M037P1_Z  GS---_6R                      0.02  AVEINV_R                      0.23 
    M037P1_Z  Obj                       0.095521  GP---_0R                         1 
    M037P1_Z  VOLM_8_R                    0.3003  VOLM_9_R                    0.1251 
    M037P1_Z  VOLM16_R                     0.181 
    M037P1_[  VOLM17_R                     0.194  VOLM_9_R                    0.1251 
    M037P1_[  R037_TP2                         1  GS---_5R                   0.06667 
    M037P1_[  VOLM18_R                     0.201  VOLM_8_R                    0.3003 
    M037P1_[  GP---_0R                         1  VOLM13_R                     0.057 
    M037P1_[  LTSY_R                     0.05022  Obj                        0.09502 
    M037P1_[  AVEINV_R                   0.24833  GS---_6R                      0.02 
    M037P1_[  INVEN_R                      0.142 
    M037P2_1  VOLM11_R                    0.0686  VOLM_7_R                      0.05 
    M037P2_1  VOLM20_R                    0.0686  VOLM_9_R                     0.096 
    M037P2_1  GS---_6R                      0.04  R037_TP2                         1 
    M037P2_1  VOLM_4_R                    0.1206  Obj                        0.54395 
    M037P2_1  GS---_3R                      0.05  INVEN_R                      0.343 
    M037P2_1  GS---_2R                      0.05  VOLM18_R                     0.096 
    M037P2_1  VOLM14_R                    0.1173  VOLM16_R                      0.05 
    M037P2_1  GP---_0R                         1  LTSY_R                     0.05782 
    M037P2_1  VOLM12_R                    0.1885  VOLM_3_R                    0.2723 
    M037P2_1  AVEINV_R                    0.2388 
    M037P2_2  LTSY_R                     0.05342  AVEINV_R                   0.25374 
    M037P2_2  VOLM_7_R                      0.05  VOLM_4_R                    0.1206 
    M037P2_2  VOLM12_R                    0.0754  GS---_2R                      0.05 
    M037P2_2  INVEN_R                      0.339  GS---_3R                      0.05 
    M037P2_2  Obj                        0.53781  GP---_0R                         1 
    M037P2_2  VOLM_3_R                    0.2723  R037_TP2                         1 
    M037P2_2  GS---_6R                      0.03  VOLM15_R                    0.1173 
    M037P2_2  VOLM19_R                     0.096  VOLM_9_R                     0.096 
    M037P2_2  VOLM13_R                    0.1955  VOLM17_R                      0.05 
    M037P2_3  VOLM17_R                      0.05  VOLM19_R                     0.096 
    M037P2_3  AVEINV_R                    0.2388  R037_TP2                         1 
    M037P2_3  VOLM_4_R                    0.2814  VOLM13_R                    0.1885 
    M037P2_3  LTSY_R                     0.05782  VOLM10_R                     0.096 
    M037P2_3  VOLM12_R                    0.0686  INVEN_R                      0.339 
    M037P2_3  VOLM_5_R                    0.1269  GS---_6R                      0.03 
    M037P2_3  Obj                        0.40418  GP---_0R                         1 
    M037P2_3  GS---_3R                       0.1  VOLM15_R                    0.1173 
    M037P2_3  VOLM_8_R                      0.05 
    M037P2_4  VOLM18_R                      0.05  GP---_0R                         1 
    M037P2_4  Obj                        0.40011  GS---_6R                      0.03 
    M037P2_4  VOLM_5_R                    0.1269  VOLM20_R                     0.096 
    M037P2_4  GS---_3R                       0.1  VOLM13_R                    0.0754 
    M037P2_4  VOLM_4_R                    0.2814  VOLM14_R                    0.1955 
    M037P2_4  R037_TP2                         1  INVEN_R                      0.317 
    M037P2_4  VOLM10_R                     0.096  AVEINV_R                   0.25374 
    M037P2_4  LTSY_R                     0.05342  VOLM16_R                    0.1173 
    M037P2_4  VOLM_8_R                      0.05 
    M037P2_5  GS---_6R                      0.03  VOLM_9_R                      0.05 
    M037P2_5  VOLM_5_R                    0.2961  GS---_4R                      0.05 
    M037P2_5  GS---_3R                      0.05  VOLM20_R                     0.096 
    M037P2_5  VOLM16_R                    0.1173  VOLM18_R                      0.05 
    M037P2_5  Obj                        0.29368  VOLM_6_R                     0.132 
    M037P2_5  VOLM13_R                    0.0686  VOLM14_R                    0.1885 
    M037P2_5  GP---_0R                         1  R037_TP2                         1 
    M037P2_5  VOLM11_R                     0.096  AVEINV_R                    0.2388 
    M037P2_5  INVEN_R                      0.317  LTSY_R                     0.05782 
    M037P2_6  INVEN_R                      0.244  VOLM19_R                      0.05 
    M037P2_6  VOLM11_R                     0.096  AVEINV_R                   0.25374 
    M037P2_6  GS---_6R                      0.03  VOLM_5_R                    0.2961 
    M037P2_6  VOLM15_R                    0.1955  R037_TP2                         1 
    M037P2_6  GS---_4R                      0.05  Obj                        0.29073 
    M037P2_6  GS---_3R                      0.05  VOLM17_R                    0.1173 
    M037P2_6  GP---_0R                         1  LTSY_R                     0.05342 
    M037P2_6  VOLM_9_R                      0.05  VOLM_6_R                     0.132 
    M037P2_6  VOLM14_R                    0.0754 
    M037P2_7  LTSY_R                     0.05782  VOLM17_R                    0.1173 
    M037P2_7  VOLM_6_R                     0.308  GP---_0R                         1 
    M037P2_7  VOLM12_R                     0.096  Obj                        0.20752 
    M037P2_7  VOLM_7_R                    0.1311  VOLM19_R                      0.05 
    M037P2_7  INVEN_R                      0.244  R037_TP2                         1 
    M037P2_7  VOLM10_R                      0.05  VOLM15_R                    0.1885 
    M037P2_7  GS---_6R                      0.03  AVEINV_R                    0.2388 
    M037P2_7  GS---_4R                       0.1  VOLM14_R                    0.0686 
    M037P2_8  VOLM20_R                      0.05  VOLM_7_R                    0.1311 
    M037P2_8  LTSY_R                     0.05342  GS---_6R                      0.03 
    M037P2_8  GP---_0R                         1  VOLM12_R                     0.096 
    M037P2_8  Obj                         0.2061  VOLM10_R                      0.05 
    M037P2_8  AVEINV_R                   0.25374  R037_TP2                         1 
    M037P2_8  VOLM_6_R                     0.308  VOLM18_R                    0.1173 
    M037P2_8  VOLM16_R                    0.1955  INVEN_R                      0.206 
    M037P2_8  GS---_4R                       0.1  VOLM15_R                    0.0754 
    M037P2_9  VOLM15_R                    0.0686  VOLM18_R                    0.1173 
    M037P2_9  GP---_0R                         1  VOLM_7_R                    0.3059 
    M037P2_9  VOLM13_R                     0.096  R037_TP2                         1 
    M037P2_9  VOLM11_R                      0.05  VOLM_8_R                    0.1287 
    M037P2_9  Obj                        0.14802  VOLM16_R                    0.1885 
    M037P2_9  INVEN_R                      0.206  GS---_6R                      0.03 
    M037P2_9  LTSY_R                     0.05782  GS---_5R                   0.03333 
    M037P2_9  VOLM20_R                      0.05  GS---_4R                      0.05 
    M037P2_9  AVEINV_R                    0.2388 
    M037P2_A  AVEINV_R                   0.25374  Obj                        0.14688 
    M037P2_A  GP---_0R                         1  GS---_5R                   0.03333 
    M037P2_A  GS---_4R                      0.05  LTSY_R                     0.05342 
    M037P2_A  VOLM_7_R                    0.3059  VOLM13_R                     0.096 
    M037P2_A  GS---_6R                      0.03  VOLM17_R                    0.1955 
    M037P2_A  R037_TP2                         1  VOLM11_R                      0.05 
    M037P2_A  INVEN_R                      0.142  VOLM_8_R                    0.1287 
    M037P2_A  VOLM19_R                    0.1173  VOLM16_R                    0.0754 
    M037P2_B  VOLM16_R                    0.0686  GS---_5R                   0.06667 
    M037P2_B  INVEN_R                      0.142  Obj                       0.097746 
    M037P2_B  VOLM_8_R                    0.3003  VOLM19_R                    0.1173 
    M037P2_B  R037_TP2                         1  GS---_6R                      0.03 
    M037P2_B  VOLM17_R                    0.1885  VOLM14_R                     0.096 
    M037P2_B  LTSY_R                     0.05782  GP---_0R                         1 
    M037P2_B  AVEINV_R                    0.2388  VOLM_9_R                    0.1251 
    M037P2_B  VOLM12_R                      0.05 
    M037P2_C  GP---_0R                         1  VOLM_8_R                    0.3003 
    M037P2_C  VOLM14_R                     0.096  VOLM12_R                      0.05 
    M037P2_C  LTSY_R                     0.05342  VOLM_9_R                    0.1251 
    M037P2_C  VOLM20_R                    0.1173  R037_TP2                         1 
    M037P2_C  AVEINV_R                   0.25374  VOLM17_R                    0.0754 
    M037P2_C  VOLM18_R                    0.1955  GS---_6R                      0.03 
    M037P2_C  Obj                       0.096663  INVEN_R                     0.1393 
    M037P2_C  GS---_5R                   0.06667 
    M048MN_1  R048_MN1                         1  Obj                     -0.0012632 
    M048RD_1  R048_RD1                         1  Obj                     -0.0010105 
    T048TM12  Obj                              0  R048_TM1                         1 
    T048TM12  R048_TM2                        -1 
    T048TM23  R048_TM3                        -1  R048_TM2                         1 
    T048TM23  Obj                              0 
    T048TM34  R048_TM3                         1  R048_TM4                        -1 
    T048TM34  Obj                              0 
    T048TM45  Obj                              0  R048_TM5                        -1 
    T048TM45  R048_TM4                         1 
    M048TF_1  AVEINV_R                     0.169  INVEN_R                      0.058 
    M048TF_1  Obj                        0.37542  VOLM17_R                     0.319 
    M048TF_1  GS+++_1R                         1  R048_TM1                         1 
    M048TF_1  LTSY_R                     0.03987  VOLM_9_R                     0.319 
    M048TF_1  GS+++_9R                         1  VOLM_1_R                     0.303 
    M048TF_1  A$___1_1                   0.01994  GP+++_0R                         1 
    M048TF_2  VOLM19_R                     0.326  A$___1_1                   0.01994 
    M048TF_2  VOLM_1_R                     0.303  GS+++10R                         1 
    M048TF_2  VOLM10_R                     0.326  LTSY_R                     0.03622 
    M048TF_2  AVEINV_R                   0.18644  Obj                        0.37053 
    M048TF_2  R048_TM1                         1  GS+++_1R                         1 
    M048TF_2  GP+++_0R                         1 
    M048TF_3  GP+++_0R                         1  GS+++_2R                         1 
    M048TF_3  VOLM18_R                     0.319  A$___1_2                   0.02448 
    M048TF_3  Obj                        0.26168  AVEINV_R                     0.169 
    M048TF_3  R048_TM2                         1  GS+++10R                         1 
    M048TF_3  VOLM10_R                     0.319  VOLM_2_R                     0.312 
    M048TF_3  LTSY_R                     0.03987  INVEN_R                      0.016 
    M048TF_4  GP+++_0R                         1  A$___1_2                   0.02448 
    M048TF_4  GS+++_2R                         1  Obj                        0.25837 
    M048TF_4  AVEINV_R                   0.18644  VOLM20_R                     0.326 
    M048TF_4  VOLM11_R                     0.326  R048_TM2                         1 
    M048TF_4  GS+++11R                         1  VOLM_2_R                     0.312 
    M048TF_4  LTSY_R                     0.03622  INVEN_R                      0.326 
    M048TF_5  VOLM11_R                     0.319  VOLM19_R                     0.319 
    M048TF_5  R048_TM3                         1  VOLM_3_R                     0.297 
    M048TF_5  GS+++11R                         1  GP+++_0R                         1 
    M048TF_5  AVEINV_R                     0.169  Obj                        0.15874 
    M048TF_5  GS+++_3R                         1  LTSY_R                     0.03987 
    M048TF_6  VOLM12_R                     0.326  VOLM_3_R                     0.297 
    M048TF_6  GS+++12R                         1  R048_TM3                         1 
    M048TF_6  GS+++_3R                         1  AVEINV_R                   0.18644 
    M048TF_6  INVEN_R                      0.319  LTSY_R                     0.03622 
    M048TF_6  Obj                        0.15578  GP+++_0R                         1 
    M048TF_7  GS+++_4R                         1  GP+++_0R                         1 
    M048TF_7  Obj                        0.10279  INVEN_R                      0.319 
    M048TF_7  VOLM12_R                     0.319  VOLM20_R                     0.319 
    M048TF_7  R048_TM4                         1  AVEINV_R                     0.169 
    M048TF_7  VOLM_4_R                     0.288  LTSY_R                     0.03987 
    M048TF_7  GS+++12R                         1 
    M048TF_8  Obj                        0.10057  GS+++_4R                         1 
    M048TF_8  VOLM13_R                     0.326  GP+++_0R                         1 
    M048TF_8  INVEN_R                        0.3  AVEINV_R                   0.18644 
    M048TF_8  R048_TM4                         1  VOLM_4_R                     0.288 
    M048TF_8  LTSY_R                     0.03622  GS+++13R                         1 
    M048TF_9  AVEINV_R                     0.169  GS+++_5R                         1 
    M048TF_9  GS+++13R                         1  GP+++_0R                         1 
    M048TF_9  VOLM13_R                     0.319  LTSY_R                     0.03987 
    M048TF_9  R048_TM5                         1  INVEN_R                        0.3 
    M048TF_9  Obj                       0.067389  VOLM_5_R                     0.279 
    M048TF_A  GS+++_5R                         1  GP+++_0R                         1 
    M048TF_A  VOLM14_R                     0.326  INVEN_R                      0.267 
    M048TF_A  R048_TM5                         1  GS+++14R                         1 
    M048TF_A  VOLM_5_R                     0.279  AVEINV_R                   0.18644 
    M048TF_A  LTSY_R                     0.03622  Obj                       0.066501 
    M048TF_B  VOLM_6_R                      0.27  LTSY_R                     0.03987 
    M048TF_B  Obj                       0.059138  GP+++_0R                         1 
    M048TF_B  R048_TM5                         1  GS+++14R                         1 
    M048TF_B  VOLM14_R                     0.319  GS+++_6R                         1 
    M048TF_B  AVEINV_R                     0.169  INVEN_R                      0.267 
    M048TF_C  AVEINV_R                   0.18644  LTSY_R                     0.03622 
    M048TF_C  VOLM15_R                     0.326  Obj                        0.05849 
    M048TF_C  VOLM_6_R                      0.27  GP+++_0R                         1 
    M048TF_C  INVEN_R                       0.22  GS+++15R                         1 
    M048TF_C  R048_TM5                         1  GS+++_6R                         1 
    M048TF_D  R048_TM5                         1  VOLM15_R                     0.319 
    M048TF_D  GS+++15R                         1  VOLM_7_R                     0.243 
    M048TF_D  GP+++_0R                         1  GS+++_7R                         1 
    M048TF_D  LTSY_R                     0.03987  Obj                       0.030706 
    M048TF_D  AVEINV_R                     0.169  INVEN_R                       0.22 
    M048TF_E  GS+++_7R                         1  INVEN_R                      0.172 
    M048TF_E  AVEINV_R                   0.18644  VOLM_7_R                     0.243 
    M048TF_E  LTSY_R                     0.03622  VOLM16_R                     0.326 
    M048TF_E  R048_TM5                         1  Obj                       0.030345 
    M048TF_E  GP+++_0R                         1 
    T048TP12  R048_TP2                        -1  Obj                              0 
    T048TP12  R048_TP1                         1 
    T048TP23  Obj                              0  R048_TP2                         1 
    T048TP23  R048_TP3                        -1 
    T048TP34  Obj                              0  R048_TP3                         1 
    T048TP34  R048_TP4                        -1 
    M048PF_1  LTSY_R                     0.04009  VOLM10_R                    0.0815 
    M048PF_1  VOLM_2_R                    0.0624  INVEN_R                      0.058 
    M048PF_1  VOLM18_R                    0.0815  GS---_1R                       0.1 
    M048PF_1  Obj                        0.63172  GS---_2R                      0.05 
    M048PF_1  GS---_5R                   0.06667  VOLM_9_R                   0.23925 
    M048PF_1  R048_TP1                         1  VOLM_1_R                    0.2424 
    M048PF_1  AVEINV_R                   0.17919  GS---_6R                      0.02 
    M048PF_1  VOLM17_R                   0.23925  GP---_0R                         1 
    M048PF_2  GP---_0R                         1  R048_TP1                         1 
    M048PF_2  GS---_1R                       0.1  VOLM10_R                    0.2445 
    M048PF_2  VOLM11_R                    0.0815  GS---_6R                      0.03 
    M048PF_2  Obj                        0.62524  AVEINV_R                    0.1955 
    M048PF_2  GS---_5R                   0.03333  INVEN_R                    0.20375 
    M048PF_2  VOLM19_R                    0.2445  LTSY_R                     0.03622 
    M048PF_2  VOLM_2_R                    0.0624  VOLM20_R                    0.0815 
    M048PF_2  GS---_2R                      0.05  VOLM_1_R                    0.2424 
    M048PF_3  R048_TP2                         1  AVEINV_R                   0.17919 
    M048PF_3  GP---_0R                         1  GS---_6R                      0.03 
    M048PF_3  VOLM18_R                   0.23925  Obj                         0.4619 
    M048PF_3  VOLM19_R                    0.0815  VOLM10_R                   0.23925 
    M048PF_3  VOLM11_R                    0.0815  GS---_5R                   0.03333 
    M048PF_3  GS---_2R                       0.1  INVEN_R                      0.016 
    M048PF_3  VOLM_3_R                    0.0594  LTSY_R                     0.04009 
    M048PF_3  VOLM_2_R                    0.2496 
    M048PF_4  GP---_0R                         1  VOLM20_R                    0.2445 
    M048PF_4  INVEN_R                      0.326  VOLM_2_R                    0.2496 
    M048PF_4  GS---_2R                       0.1  VOLM11_R                    0.2445 
    M048PF_4  GS---_6R                      0.03  LTSY_R                     0.03622 
    M048PF_4  Obj                        0.45761  R048_TP2                         1 
    M048PF_4  VOLM12_R                    0.0815  AVEINV_R                    0.1955 
    M048PF_4  VOLM_3_R                    0.0594 
    M048PF_5  VOLM11_R                   0.23925  VOLM12_R                    0.0815 
    M048PF_5  VOLM19_R                   0.23925  VOLM_3_R                    0.2376 
    M048PF_5  INVEN_R                    0.20375  LTSY_R                     0.04009 
    M048PF_5  Obj                        0.30731  VOLM20_R                    0.0815 
    M048PF_5  R048_TP3                         1  GS---_2R                      0.05 
    M048PF_5  AVEINV_R                   0.17919  GP---_0R                         1 
    M048PF_5  GS---_3R                      0.05  VOLM_4_R                    0.0576 
    M048PF_5  GS---_6R                      0.04 
    M048PF_6  GS---_6R                      0.02  INVEN_R                      0.319 
    M048PF_6  VOLM_4_R                    0.0576  VOLM13_R                    0.0815 
    M048PF_6  GS---_2R                      0.05  VOLM12_R                    0.2445 
    M048PF_6  R048_TP3                         1  GP---_0R                         1 
    M048PF_6  GS---_3R                      0.05  AVEINV_R                    0.1955 
    M048PF_6  LTSY_R                     0.03622  Obj                        0.30384 
    M048PF_6  VOLM_3_R                    0.2376 
    M048PF_7  VOLM13_R                    0.0815  R048_TP4                         1 
    M048PF_7  LTSY_R                     0.04009  GS---_6R                      0.03 
    M048PF_7  GS---_3R                       0.1  VOLM_5_R                    0.0558 
    M048PF_7  AVEINV_R                   0.17919  VOLM12_R                   0.23925 
    M048PF_7  VOLM20_R                   0.23925  GP---_0R                         1 
    M048PF_7  Obj                        0.20376  INVEN_R                      0.319 
    M048PF_7  VOLM_4_R                    0.2304 
    M048PF_8  VOLM_5_R                    0.0558  INVEN_R                        0.3 
    M048PF_8  GP---_0R                         1  Obj                        0.20137 
    M048PF_8  VOLM14_R                    0.0815  GS---_6R                      0.02 
    M048PF_8  AVEINV_R                    0.1955  LTSY_R                     0.03622 
    M048PF_8  VOLM13_R                    0.2445  R048_TP4                         1 
    M048PF_8  VOLM_4_R                    0.2304  GS---_3R                       0.1 
    M048PF_9  GS---_6R                      0.02  GS---_4R                      0.05 
    M048PF_9  R048_TP4                         1  VOLM_6_R                     0.054 
    M048PF_9  VOLM13_R                   0.23925  VOLM_5_R                    0.2232 
    M048PF_9  GS---_3R                      0.05  LTSY_R                     0.04009 
    M048PF_9  VOLM14_R                    0.0815  AVEINV_R                   0.17919 
    M048PF_9  INVEN_R                        0.3  GP---_0R                         1 
    M048PF_9  Obj                        0.13887 
    M048PF_A  GS---_6R                      0.02  VOLM14_R                    0.2445 
    M048PF_A  R048_TP4                         1  Obj                        0.13774 
    M048PF_A  VOLM15_R                    0.0815  GS---_3R                      0.05 
    M048PF_A  AVEINV_R                    0.1955  LTSY_R                     0.03622 
    M048PF_A  VOLM_6_R                     0.054  INVEN_R                      0.267 
    M048PF_A  GP---_0R                         1  GS---_4R                      0.05 
    M048PF_A  VOLM_5_R                    0.2232 
    M048PF_B  VOLM_6_R                     0.216  GS---_6R                      0.02 
    M048PF_B  VOLM_7_R                    0.0486  VOLM15_R                    0.0815 
    M048PF_B  GS---_4R                       0.1  AVEINV_R                   0.17919 
    M048PF_B  VOLM14_R                   0.23925  R048_TP4                         1 
    M048PF_B  GP---_0R                         1  Obj                        0.10302 
    M048PF_B  INVEN_R                      0.267  LTSY_R                     0.04009 
    M048PF_C  LTSY_R                     0.03622  INVEN_R                       0.22 
    M048PF_C  VOLM_6_R                     0.216  VOLM_7_R                    0.0486 
    M048PF_C  Obj                        0.10221  GP---_0R                         1 
    M048PF_C  GS---_4R                       0.1  GS---_6R                      0.02 
    M048PF_C  R048_TP4                         1  VOLM16_R                    0.0815 
    M048PF_C  AVEINV_R                    0.1955  VOLM15_R                    0.2445 
    M048PF_D  GS---_5R                   0.03333  GP---_0R                         1 
    M048PF_D  INVEN_R                       0.22  VOLM_8_R                    0.0462 
    M048PF_D  Obj                       0.061713  VOLM_7_R                    0.1944 
    M048PF_D  GS---_4R                      0.05  R048_TP4                         1 
    M048PF_D  AVEINV_R                   0.17919  LTSY_R                     0.04009 
    M048PF_D  GS---_6R                      0.02  VOLM15_R                   0.23925 
    M048PF_D  VOLM16_R                    0.0815 
    M048PF_E  GS---_4R                      0.05  Obj                         0.0611 
    M048PF_E  GS---_5R                   0.03333  GS---_6R                      0.02 
    M048PF_E  LTSY_R                     0.03622  R048_TP4                         1 
    M048PF_E  VOLM17_R                    0.0815  AVEINV_R                    0.1955 
    M048PF_E  VOLM_7_R                    0.1944  VOLM16_R                    0.2445 
    M048PF_E  VOLM_8_R                    0.0462  GP---_0R                         1 
    M048PF_E  INVEN_R                      0.172 
    M052MN_1  R052_MN1                         1  Obj                     -0.0012632 
    M052RD_1  Obj                     -0.0010105  R052_RD1                         1 
    T052TM12  Obj                              0  R052_TM2                        -1 
    T052TM12  R052_TM1                         1 
    T052TM23  R052_TM2                         1  Obj                              0 
    T052TM23  R052_TM3                        -1 
    T052TM34  R052_TM3                         1  Obj                              0 
    T052TM34  R052_TM4                        -1 
    T052TM45  Obj                              0  R052_TM5                        -1 
    T052TM45  R052_TM4                         1 
    M052TF_1  R052_TM1                         1  GP+++_0R                         1 
    M052TF_1  LTSY_R                     0.05288  GS+++_9R                         1 
    M052TF_1  A$___1_1                   0.01994  VOLM_1_R                     0.278 
    M052TF_1  VOLM_9_R                     0.423  Obj                       0.075315 
    M052TF_1  INVEN_R                      0.076  GS+++_1R                         1 
    M052TF_1  VOLM17_R                     0.423  AVEINV_R                       0.2 
    M052TF_2  LTSY_R                     0.05311  A$___1_1                   0.01994 
    M052TF_2  VOLM10_R                     0.478  GS+++_1R                         1 
    M052TF_2  VOLM19_R                     0.478  R052_TM1                         1 
    M052TF_2  AVEINV_R                   0.23089  GP+++_0R                         1 
    M052TF_2  GS+++10R                         1  Obj                       0.060634 
    M052TF_2  VOLM_1_R                     0.278 
    M052TF_3  AVEINV_R                       0.2  R052_TM2                         1 
    M052TF_3  GS+++10R                         1  Obj                        0.13185 
    M052TF_3  GP+++_0R                         1  VOLM_2_R                     0.302 
    M052TF_3  INVEN_R                      0.034  VOLM18_R                     0.423 
    M052TF_3  A$___1_2                   0.02448  LTSY_R                     0.05288 
    M052TF_3  GS+++_2R                         1  VOLM10_R                     0.423 
    M052TF_4  LTSY_R                     0.05311  GS+++_2R                         1 
    M052TF_4  R052_TM2                         1  VOLM20_R                     0.478 
    M052TF_4  VOLM11_R                     0.478  A$___1_2                   0.02448 
    M052TF_4  GS+++11R                         1  AVEINV_R                   0.23089 
    M052TF_4  INVEN_R                      0.478  GP+++_0R                         1 
    M052TF_4  VOLM_2_R                     0.302  Obj                        0.12194 
    M052TF_5  VOLM11_R                     0.423  LTSY_R                     0.05288 
    M052TF_5  AVEINV_R                       0.2  R052_TM3                         1 
    M052TF_5  VOLM_3_R                     0.327  GS+++11R                         1 
    M052TF_5  GS+++_3R                         1  GP+++_0R                         1 
    M052TF_5  VOLM19_R                     0.423  Obj                        0.09368 
    M052TF_6  Obj                       0.086181  R052_TM3                         1 
    M052TF_6  GP+++_0R                         1  VOLM12_R                     0.478 
    M052TF_6  GS+++12R                         1  LTSY_R                     0.05311 
    M052TF_6  VOLM_3_R                     0.327  GS+++_3R                         1 
    M052TF_6  INVEN_R                      0.423  AVEINV_R                   0.23089 
    M052TF_7  Obj                       0.068434  VOLM12_R                     0.423 
    M052TF_7  LTSY_R                     0.05288  GP+++_0R                         1 
    M052TF_7  GS+++12R                         1  R052_TM4                         1 
    M052TF_7  VOLM20_R                     0.423  INVEN_R                      0.423 
    M052TF_7  GS+++_4R                         1  AVEINV_R                       0.2 
    M052TF_7  VOLM_4_R                     0.341 
    M052TF_8  VOLM_4_R                     0.341  Obj                       0.063485 
    M052TF_8  INVEN_R                      0.345  LTSY_R                     0.05311 
    M052TF_8  GS+++13R                         1  GS+++_4R                         1 
    M052TF_8  R052_TM4                         1  AVEINV_R                   0.23089 
    M052TF_8  GP+++_0R                         1  VOLM13_R                     0.478 
    M052TF_9  Obj                       0.044678  GS+++13R                         1 
    M052TF_9  VOLM_5_R                     0.343  INVEN_R                      0.345 
    M052TF_9  R052_TM5                         1  GP+++_0R                         1 
    M052TF_9  GS+++_5R                         1  VOLM13_R                     0.423 
    M052TF_9  AVEINV_R                       0.2  LTSY_R                     0.05288 
    M052TF_A  GP+++_0R                         1  INVEN_R                      0.299 
    M052TF_A  AVEINV_R                   0.23089  R052_TM5                         1 
    M052TF_A  VOLM14_R                     0.478  VOLM_5_R                     0.343 
    M052TF_A  GS+++_5R                         1  GS+++14R                         1 
    M052TF_A  Obj                       0.041872  LTSY_R                     0.05311 
    M052TF_B  INVEN_R                      0.299  LTSY_R                     0.05288 
    M052TF_B  VOLM_6_R                     0.331  AVEINV_R                       0.2 
    M052TF_B  GS+++14R                         1  Obj                       0.026533 
    M052TF_B  VOLM14_R                     0.423  GP+++_0R                         1 
    M052TF_B  GS+++_6R                         1  R052_TM5                         1 
    M052TF_C  Obj                       0.024657  LTSY_R                     0.05311 
    M052TF_C  VOLM_6_R                     0.331  GS+++15R                         1 
    M052TF_C  GS+++_6R                         1  GP+++_0R                         1 
    M052TF_C  VOLM15_R                     0.478  R052_TM5                         1 
    M052TF_C  INVEN_R                      0.245  AVEINV_R                   0.23089 
    M052TF_D  INVEN_R                      0.245  GP+++_0R                         1 
    M052TF_D  R052_TM5                         1  VOLM15_R                     0.423 
    M052TF_D  GS+++_7R                         1  LTSY_R                     0.05288 
    M052TF_D  GS+++15R                         1  AVEINV_R                       0.2 
    M052TF_D  VOLM_7_R                     0.329  Obj                       0.016863 
    M052TF_E  VOLM_7_R                     0.329  R052_TM5                         1 
    M052TF_E  GP+++_0R                         1  Obj                       0.015622 
    M052TF_E  LTSY_R                     0.05311  VOLM16_R                     0.478 
    M052TF_E  INVEN_R                      0.178  AVEINV_R                   0.23089 
    M052TF_E  GS+++_7R                         1 
    M083MN_1  R083_MN1                         1  Obj                     -0.0012632 
    M083MN21  Obj                     -0.0012632  R083_GM2                         1 
    M083RD_1  Obj                     -0.0010105  R083_RD1                         1 
    M083GB_1  R083_GR2                         1  Obj                       0.049699 
    M083GB21  Obj                       0.049699  R083_GM2                         1 
    M092MN_1  Obj                     -0.0012632  R092_MN2                         1 
    M092RD_1  R092_RD1                         1  Obj                     -0.0010105 
RHS
    RHS       LC123                      7392000  DEDO3_1R                         0 
    RHS       DEDO3_2R                         0  DEDO3_3R                         0 
    RHS       DEDO3_4R                         0  DEDO3_5R                         0 
    RHS       DEDO3_6R                         0  DEDO3_7R                         0 
    RHS       DEDO3_8R                         0  DEDO3_9R                         0 
    RHS       DEDO310R                         0  DEDO311R                         0 
    RHS       DEDO312R                         0  DEDO313R                         0 
    RHS       DEDO314R                         0  DEDO315R                         0 
    RHS       DEDO5_1R                         0  DEDO5_2R                         0 
    RHS       DEDO5_3R                         0  BR___1_1                      2345 
    RHS       BR___2_2                      2800  BR___2_3                      2800 
    RHS       VOLM_1_R                         0  VOLM_2_R                         0 
    RHS       VOLM_3_R                         0  VOLM_4_R                         0 
    RHS       VOLM_5_R                         0  VOLM_6_R                         0 
    RHS       VOLM_7_R                         0  VOLM_8_R                         0 
    RHS       VOLM_9_R                         0  VOLM10_R                         0 
    RHS       VOLM11_R                         0  VOLM12_R                         0 
    RHS       VOLM13_R                         0  VOLM14_R                         0 
    RHS       VOLM15_R                         0  VOLM16_R                         0 
    RHS       VOLM17_R                         0  VOLM18_R                         0 
    RHS       VOLM19_R                         0  VOLM20_R                         0 
    RHS       BHVG_2                           0  BHVL_2                           0 
    RHS       BHVG_3                           0  BHVL_3                           0 
    RHS       BHVG_4                           0  BHVL_4                           0 
    RHS       BHVG_5                           0  BHVL_5                           0 
    RHS       BHVG_6                           0  BHVL_6                           0 
    RHS       BHVG_7                           0  BHVL_7                           0 
    RHS       BHVG_8                           0  BHVL_8                           0 
    RHS       BHVG_9                           0  BHVL_9                           0 
    RHS       BHVG10                           0  BHVL10                           0 
    RHS       BHVG11                           0  BHVL11                           0 
    RHS       BHVG12                           0  BHVL12                           0 
    RHS       BHVG13                           0  BHVL13                           0 
    RHS       BHVG14                           0  BHVL14                           0 
    RHS       BHVG15                           0  BHVL15                           0 
    RHS       BHVG16                           0  BHVL16                           0 
    RHS       BHVG17                           0  BHVL17                           0 
    RHS       BHVG18                           0  BHVL18                           0 
    RHS       BHVG19                           0  BHVL19                           0 
    RHS       BHVG20                           0  BHVL20                           0 
    RHS       SYNDY                            0  LTSY_R                           0 
    RHS       LTSYCT                      285000  AVEINV_R                         0 
    RHS       ENDINVCT                         0  INVEN_R                          0 
    RHS       A$___1_1                      3500  A$___1_2                      3500 
    RHS       A$_4-8_1                      4712  A$_4-8_2                      4712 
    RHS       A$_4-8_3                      4712  A$_4-8_4                      4712 
    RHS       GP+++_0R                         0  GS+++_1R                         0 
    RHS       GS+++_2R                         0  GS+++_3R                         0 
    RHS       GS+++_4R                         0  GS+++_5R                         0 
    RHS       GS+++_6R                         0  GS+++_7R                         0 
    RHS       GS+++_8R                         0  GS+++_9R                         0 
    RHS       GS+++10R                         0  GS+++11R                         0 
    RHS       GS+++12R                         0  GS+++13R                         0 
    RHS       GS+++14R                         0  GS+++15R                         0 
    RHS       GP---_0R                         0  GS---_1R                         0 
    RHS       GS---_2R                         0  GS---_3R                         0 
    RHS       GS---_4R                         0  GS---_5R                         0 
    RHS       GS---_6R                         0  R012_MN1                         0 
    RHS       R012_RD1                         0  R012_TM1                         0 
    RHS       R012_TM2                         0  R012_TM3                         0 
    RHS       R012_TM4                         0  R012_TM5                         0 
    RHS       R012_TM6                         0  R012_TP1                         0 
    RHS       R012_TP2                         0  R012_TP3                         0 
    RHS       R012_TP4                         0  R012_TP5                         0 
    RHS       R012_TP6                         0  R037_MN1                         0 
    RHS       R037_RD1                         0  R037_TM2                         0 
    RHS       R037_TP2                         0  R048_MN1                         0 
    RHS       R048_RD1                         0  R048_TM1                         0 
    RHS       R048_TM2                         0  R048_TM3                         0 
    RHS       R048_TM4                         0  R048_TM5                         0 
    RHS       R048_TP1                         0  R048_TP2                         0 
    RHS       R048_TP3                         0  R048_TP4                         0 
    RHS       R052_MN1                         0  R052_RD1                         0 
    RHS       R052_TM1                         0  R052_TM2                         0 
    RHS       R052_TM3                         0  R052_TM4                         0 
    RHS       R052_TM5                         0  R083_MN1                         0 
    RHS       R083_GM2                         0  R083_RD1                         0 
    RHS       R083_GR2                         0  R092_MN2                         0 
    RHS       R092_RD1                         0  AZ__20                        2640 
    RHS       AZ__80                        2800  AZ__90                        2640 
    RHS       AZ_100                        2800 
RANGES
    RANGE     LTSYCT                      284990 
BOUNDS
 UP Bound     DEDO3_11                    200000 
 PL Bound     DEDO3_12                           
 UP Bound     DEDO3_21                    220000 
 PL Bound     DEDO3_22                           
 UP Bound     DEDO3_31                    275000 
 PL Bound     DEDO3_32                           
 UP Bound     DEDO3_41                    275000 
 PL Bound     DEDO3_42                           
 UP Bound     DEDO3_51                    298000 
 PL Bound     DEDO3_52                           
 UP Bound     DEDO3_61                    298000 
 PL Bound     DEDO3_62                           
 UP Bound     DEDO3_71                    298000 
 PL Bound     DEDO3_72                           
 UP Bound     DEDO3_81                    298000 
 PL Bound     DEDO3_82                           
 UP Bound     DEDO3_91                    298000 
 PL Bound     DEDO3_92                           
 UP Bound     DEDO3101                    298000 
 PL Bound     DEDO3102                           
 UP Bound     DEDO3111                    298000 
 PL Bound     DEDO3112                           
 UP Bound     DEDO3121                    298000 
 PL Bound     DEDO3122                           
 UP Bound     DEDO3131                    298000 
 PL Bound     DEDO3132                           
 UP Bound     DEDO3141                    298000 
 PL Bound     DEDO3142                           
 UP Bound     DEDO3151                    298000 
 PL Bound     DEDO3152                           
 UP Bound     DEDO5_11                    120000 
 UP Bound     DEDO5_12                   9999999 
 UP Bound     DEDO5_21                    135000 
 UP Bound     DEDO5_22                   9999999 
 UP Bound     DEDO5_31                    147000 
 UP Bound     DEDO5_32                   9999999 
 PL Bound     VOLM_1                             
 PL Bound     VOLM_2                             
 PL Bound     VOLM_3                             
 PL Bound     VOLM_4                             
 PL Bound     VOLM_5                             
 PL Bound     VOLM_6                             
 PL Bound     VOLM_7                             
 PL Bound     VOLM_8                             
 PL Bound     VOLM_9                             
 PL Bound     VOLM10                             
 PL Bound     VOLM11                             
 PL Bound     VOLM12                             
 PL Bound     VOLM13                             
 PL Bound     VOLM14                             
 PL Bound     VOLM15                             
 PL Bound     VOLM16                             
 PL Bound     VOLM17                             
 PL Bound     VOLM18                             
 PL Bound     VOLM19                             
 PL Bound     VOLM20                             
 PL Bound     LTSY                               
 PL Bound     AVEINV                             
 PL Bound     INVEN                              
 PL Bound     GP+++_0                            
 PL Bound     GP---_0                            
 PL Bound     A___21_1                           
 FX Bound     A___22_1                      2640 
 PL Bound     A___23_1                           
 PL Bound     A___23_2                           
 PL Bound     A___81_1                           
 PL Bound     A___82_1                           
 FX Bound     A___83_1                         0 
 FX Bound     A___83_2                         0 
 PL Bound     A___84_1                           
 PL Bound     A___84_2                           
 PL Bound     A___91_1                           
 PL Bound     A___92_1                           
 PL Bound     A___93_1                           
 PL Bound     A___93_2                           
 PL Bound     A__101_1                           
 PL Bound     A__102_1                           
 PL Bound     A__103_1                           
 PL Bound     A__103_2                           
 PL Bound     A__104_1                           
 PL Bound     A__104_2                           
 PL Bound     A__105_1                           
 PL Bound     A__105_2                           
 PL Bound     M012MN_1                           
 PL Bound     M012RD_1                           
 PL Bound     T012TM12                           
 PL Bound     T012TM23                           
 PL Bound     T012TM34                           
 PL Bound     T012TM45                           
 PL Bound     T012TM56                           
 PL Bound     M012TF_1                           
 PL Bound     M012TF_2                           
 PL Bound     M012TF_3                           
 PL Bound     M012TF_4                           
 PL Bound     M012TF_5                           
 PL Bound     M012TF_6                           
 PL Bound     M012TF_7                           
 PL Bound     M012TF_8                           
 PL Bound     M012TF_9                           
 PL Bound     M012TF_A                           
 PL Bound     M012TF_B                           
 PL Bound     M012TF_C                           
 PL Bound     M012TF_D                           
 PL Bound     M012TF_E                           
 PL Bound     M012T1_1                           
 PL Bound     M012T1_2                           
 PL Bound     M012T1_3                           
 PL Bound     M012T1_4                           
 PL Bound     M012T1_5                           
 PL Bound     M012T1_6                           
 PL Bound     M012T1_7                           
 PL Bound     M012T1_8                           
 PL Bound     M012T1_9                           
 PL Bound     M012T1_A                           
 PL Bound     M012T1_B                           
 PL Bound     M012T1_C                           
 PL Bound     M012T1_D                           
 PL Bound     M012T1_E                           
 PL Bound     M012T1_F                           
 PL Bound     M012T1_G                           
 PL Bound     M012T1_H                           
 PL Bound     M012T1_I                           
 PL Bound     M012T1_J                           
 PL Bound     M012T1_K                           
 PL Bound     M012T1_L                           
 PL Bound     M012T1_M                           
 PL Bound     M012T1_N                           
 PL Bound     M012T1_O                           
 PL Bound     M012T1_P                           
 PL Bound     M012T1_Q                           
 PL Bound     M012T1_R                           
 PL Bound     M012T1_S                           
 PL Bound     M012T1_T                           
 PL Bound     M012T1_U                           
 PL Bound     M012T1_V                           
 PL Bound     M012T1_W                           
 PL Bound     M012T1_X                           
 PL Bound     M012T1_Y                           
 PL Bound     M012T1_Z                           
 PL Bound     M012T1_[                           
 PL Bound     M012T1_]                           
 PL Bound     M012T1_#                           
 PL Bound     M012T1_^                           
 PL Bound     M012T1_)                           
 PL Bound     M012T1_-                           
 PL Bound     M012T1_+                           
 PL Bound     M012T2_1                           
 PL Bound     M012T2_2                           
 PL Bound     M012T2_3                           
 PL Bound     M012T2_4                           
 PL Bound     M012T2_5                           
 PL Bound     M012T2_6                           
 PL Bound     M012T2_7                           
 PL Bound     M012T2_8                           
 PL Bound     M012T2_9                           
 PL Bound     M012T2_A                           
 PL Bound     M012T2_B                           
 PL Bound     M012T2_C                           
 PL Bound     M012T2_D                           
 PL Bound     M012T2_E                           
 PL Bound     T012TP12                           
 PL Bound     T012TP23                           
 PL Bound     T012TP34                           
 PL Bound     T012TP45                           
 PL Bound     T012TP56                           
 PL Bound     M012PF_1                           
 PL Bound     M012PF_2                           
 PL Bound     M012PF_3                           
 PL Bound     M012PF_4                           
 PL Bound     M012PF_5                           
 PL Bound     M012PF_6                           
 PL Bound     M012PF_7                           
 PL Bound     M012PF_8                           
 PL Bound     M012PF_9                           
 PL Bound     M012PF_A                           
 PL Bound     M012PF_B                           
 PL Bound     M012PF_C                           
 PL Bound     M012PF_D                           
 PL Bound     M012PF_E                           
 PL Bound     M012P1_1                           
 PL Bound     M012P1_2                           
 PL Bound     M012P1_3                           
 PL Bound     M012P1_4                           
 PL Bound     M012P1_5                           
 PL Bound     M012P1_6                           
 PL Bound     M012P1_7                           
 PL Bound     M012P1_8                           
 PL Bound     M012P1_9                           
 PL Bound     M012P1_A                           
 PL Bound     M012P1_B                           
 PL Bound     M012P1_C                           
 PL Bound     M012P1_D                           
 PL Bound     M012P1_E                           
 PL Bound     M012P1_F                           
 PL Bound     M012P1_G                           
 PL Bound     M012P1_H                           
 PL Bound     M012P1_I                           
 PL Bound     M012P1_J                           
 PL Bound     M012P1_K                           
 PL Bound     M012P1_L                           
 PL Bound     M012P1_M                           
 PL Bound     M012P1_N                           
 PL Bound     M012P1_O                           
 PL Bound     M012P1_P                           
 PL Bound     M012P1_Q                           
 PL Bound     M012P1_R                           
 PL Bound     M012P1_S                           
 PL Bound     M012P1_T                           
 PL Bound     M012P1_U                           
 PL Bound     M012P1_V                           
 PL Bound     M012P1_W                           
 PL Bound     M012P1_X                           
 PL Bound     M012P1_Y                           
 PL Bound     M012P1_Z                           
 PL Bound     M012P1_[                           
 PL Bound     M012P1_]                           
 PL Bound     M012P1_#                           
 PL Bound     M012P1_^                           
 PL Bound     M012P1_)                           
 PL Bound     M012P1_-                           
 PL Bound     M012P1_+                           
 PL Bound     M012P2_1                           
 PL Bound     M012P2_2                           
 PL Bound     M012P2_3                           
 PL Bound     M012P2_4                           
 PL Bound     M012P2_5                           
 PL Bound     M012P2_6                           
 PL Bound     M012P2_7                           
 PL Bound     M012P2_8                           
 PL Bound     M012P2_9                           
 PL Bound     M012P2_A                           
 PL Bound     M012P2_B                           
 PL Bound     M012P2_C                           
 PL Bound     M012P2_D                           
 PL Bound     M012P2_E                           
 PL Bound     M037MN_1                           
 PL Bound     M037RD_1                           
 PL Bound     M037TF_1                           
 PL Bound     M037TF_2                           
 PL Bound     M037TF_3                           
 PL Bound     M037TF_4                           
 PL Bound     M037TF_5                           
 PL Bound     M037TF_6                           
 PL Bound     M037TF_7                           
 PL Bound     M037TF_8                           
 PL Bound     M037TF_9                           
 PL Bound     M037TF_A                           
 PL Bound     M037TF_B                           
 PL Bound     M037TF_C                           
 PL Bound     M037T1_1                           
 PL Bound     M037T1_2                           
 PL Bound     M037T1_3                           
 PL Bound     M037T1_4                           
 PL Bound     M037T1_5                           
 PL Bound     M037T1_6                           
 PL Bound     M037T1_7                           
 PL Bound     M037T1_8                           
 PL Bound     M037T1_9                           
 PL Bound     M037T1_A                           
 PL Bound     M037T1_B                           
 PL Bound     M037T1_C                           
 PL Bound     M037T1_D                           
 PL Bound     M037T1_E                           
 PL Bound     M037T1_F                           
 PL Bound     M037T1_G                           
 PL Bound     M037T1_H                           
 PL Bound     M037T1_I                           
 PL Bound     M037T1_J                           
 PL Bound     M037T1_K                           
 PL Bound     M037T1_L                           
 PL Bound     M037T1_M                           
 PL Bound     M037T1_N                           
 PL Bound     M037T1_O                           
 PL Bound     M037T1_P                           
 PL Bound     M037T1_Q                           
 PL Bound     M037T1_R                           
 PL Bound     M037T1_S                           
 PL Bound     M037T1_T                           
 PL Bound     M037T1_U                           
 PL Bound     M037T1_V                           
 PL Bound     M037T1_W                           
 PL Bound     M037T1_X                           
 PL Bound     M037T1_Y                           
 PL Bound     M037T1_Z                           
 PL Bound     M037T1_[                           
 PL Bound     M037T2_1                           
 PL Bound     M037T2_2                           
 PL Bound     M037T2_3                           
 PL Bound     M037T2_4                           
 PL Bound     M037T2_5                           
 PL Bound     M037T2_6                           
 PL Bound     M037T2_7                           
 PL Bound     M037T2_8                           
 PL Bound     M037T2_9                           
 PL Bound     M037T2_A                           
 PL Bound     M037T2_B                           
 PL Bound     M037T2_C                           
 PL Bound     M037PF_1                           
 PL Bound     M037PF_2                           
 PL Bound     M037PF_3                           
 PL Bound     M037PF_4                           
 PL Bound     M037PF_5                           
 PL Bound     M037PF_6                           
 PL Bound     M037PF_7                           
 PL Bound     M037PF_8                           
 PL Bound     M037PF_9                           
 PL Bound     M037PF_A                           
 PL Bound     M037PF_B                           
 PL Bound     M037PF_C                           
 PL Bound     M037P1_1                           
 PL Bound     M037P1_2                           
 PL Bound     M037P1_3                           
 PL Bound     M037P1_4                           
 PL Bound     M037P1_5                           
 PL Bound     M037P1_6                           
 PL Bound     M037P1_7                           
 PL Bound     M037P1_8                           
 PL Bound     M037P1_9                           
 PL Bound     M037P1_A                           
 PL Bound     M037P1_B                           
 PL Bound     M037P1_C                           
 PL Bound     M037P1_D                           
 PL Bound     M037P1_E                           
 PL Bound     M037P1_F                           
 PL Bound     M037P1_G                           
 PL Bound     M037P1_H                           
 PL Bound     M037P1_I                           
 PL Bound     M037P1_J                           
 PL Bound     M037P1_K                           
 PL Bound     M037P1_L                           
 PL Bound     M037P1_M                           
 PL Bound     M037P1_N                           
 PL Bound     M037P1_O                           
 PL Bound     M037P1_P                           
 PL Bound     M037P1_Q                           
 PL Bound     M037P1_R                           
 PL Bound     M037P1_S                           
 PL Bound     M037P1_T                           
 PL Bound     M037P1_U                           
 PL Bound     M037P1_V                           
 PL Bound     M037P1_W                           
 PL Bound     M037P1_X                           
 PL Bound     M037P1_Y                           
 PL Bound     M037P1_Z                           
 PL Bound     M037P1_[                           
 PL Bound     M037P2_1                           
 PL Bound     M037P2_2                           
 PL Bound     M037P2_3                           
 PL Bound     M037P2_4                           
 PL Bound     M037P2_5                           
 PL Bound     M037P2_6                           
 PL Bound     M037P2_7                           
 PL Bound     M037P2_8                           
 PL Bound     M037P2_9                           
 PL Bound     M037P2_A                           
 PL Bound     M037P2_B                           
 PL Bound     M037P2_C                           
 PL Bound     M048MN_1                           
 PL Bound     M048RD_1                           
 PL Bound     T048TM12                           
 PL Bound     T048TM23                           
 PL Bound     T048TM34                           
 PL Bound     T048TM45                           
 PL Bound     M048TF_1                           
 PL Bound     M048TF_2                           
 PL Bound     M048TF_3                           
 PL Bound     M048TF_4                           
 PL Bound     M048TF_5                           
 PL Bound     M048TF_6                           
 PL Bound     M048TF_7                           
 PL Bound     M048TF_8                           
 PL Bound     M048TF_9                           
 PL Bound     M048TF_A                           
 PL Bound     M048TF_B                           
 PL Bound     M048TF_C                           
 PL Bound     M048TF_D                           
 PL Bound     M048TF_E                           
 PL Bound     T048TP12                           
 PL Bound     T048TP23                           
 PL Bound     T048TP34                           
 PL Bound     M048PF_1                           
 PL Bound     M048PF_2                           
 PL Bound     M048PF_3                           
 PL Bound     M048PF_4                           
 PL Bound     M048PF_5                           
 PL Bound     M048PF_6                           
 PL Bound     M048PF_7                           
 PL Bound     M048PF_8                           
 PL Bound     M048PF_9                           
 PL Bound     M048PF_A                           
 PL Bound     M048PF_B                           
 PL Bound     M048PF_C                           
 PL Bound     M048PF_D                           
 PL Bound     M048PF_E                           
 PL Bound     M052MN_1                           
 PL Bound     M052RD_1                           
 PL Bound     T052TM12                           
 PL Bound     T052TM23                           
 PL Bound     T052TM34                           
 PL Bound     T052TM45                           
 PL Bound     M052TF_1                           
 PL Bound     M052TF_2                           
 PL Bound     M052TF_3                           
 PL Bound     M052TF_4                           
 PL Bound     M052TF_5                           
 PL Bound     M052TF_6                           
 PL Bound     M052TF_7                           
 PL Bound     M052TF_8                           
 PL Bound     M052TF_9                           
 PL Bound     M052TF_A                           
 PL Bound     M052TF_B                           
 PL Bound     M052TF_C                           
 PL Bound     M052TF_D                           
 PL Bound     M052TF_E                           
 PL Bound     M083MN_1                           
 PL Bound     M083MN21                           
 PL Bound     M083RD_1                           
 PL Bound     M083GB_1                           
 PL Bound     M083GB21                           
 PL Bound     M092MN_1                           
 PL Bound     M092RD_1                           
ENDATA