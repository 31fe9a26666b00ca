!! Reference-side plugin: a mergerTreeNodeEvolver that evolves nodes on a B200 through libglcb200.so.
!!
!! To be added to the Galacticus tree as source/merger_trees/node_evolver/B200.F90 (next to standard.F90, which it
!! extends) together with integration/B200_interface.F90.  Selected by
!!
!!   <mergerTreeNodeEvolver value="B200"/>
!!
!! It cannot be compiled in the image this repository is developed in (no gfortran >= 16, GSL, HDF5, FoX:
!! SURVEY.md 8c); it is written against the reference's directive preprocessor (functionClass, inputParameter,
!! objectBuilder) and the generated treeNode / nodeComponent interfaces exactly as standard.F90 is.  Everything that
!! is not the ODE solve itself -- promote, merge, isAccurate, the parameter surface -- is inherited.

  !!{RST
  A merger tree node evolver which performs the differential evolution of nodes on an NVIDIA B200 through the
  ``libglcb200`` C-ABI library.
  !!}

  use, intrinsic :: ISO_C_Binding              , only : c_ptr                , c_null_ptr       , c_int32_t, c_int64_t, &
       &                                                c_double             , c_associated     , c_loc
  use            :: Node_Evolver_B200_Interface
  use            :: Cosmology_Parameters       , only : cosmologyParametersClass
  use            :: Cooling_Functions          , only : coolingFunctionClass
  use            :: Chemical_States            , only : chemicalStateClass
  use            :: Dark_Matter_Halo_Scales    , only : darkMatterHaloScaleClass
  use            :: Accretion_Disks            , only : accretionDisksClass

  !![
  <mergerTreeNodeEvolver name="mergerTreeNodeEvolverB200" docformat="rst">
    <description>
    A merger tree node evolver which hands the differential evolution of nodes (``standardEvolve``,
    ``standardODEs``, ``standardPostStepProcessing`` and the GSL driver chain below them) to ``libglcb200``, a
    B200-native batched Cash--Karp solver with the standard-component rate functions of ``parameters/quickTest.xml`` as
    hand-written FP64 device functions. Nodes are exchanged as records of ``GLC_NPROP`` doubles: the 24 numerically
    integrated properties in the order of ``treeNodeSerializeValuesToArray`` followed by the analytically solved
    properties and the galactic structure warm starts the rate functions read.

    ``evolve`` evolves one node (a batch of one: correct, but only useful for testing); ``evolveBatch`` is what
    :galacticus-class:`mergerTreeEvolverB200` calls with every currently evolvable node of a set of forests.
    </description>
  </mergerTreeNodeEvolver>
  !!]
  type, extends(mergerTreeNodeEvolverStandard) :: mergerTreeNodeEvolverB200
     !!{RST
     Implementation of a merger tree node evolver which evolves nodes on a B200 through ``libglcb200``.
     !!}
     private
     type   (c_ptr    ) :: evolver            =c_null_ptr
     type   (glcParams) :: parameters_
     integer            :: deviceOrdinal
     logical            :: resolveInterruptsOnDevice
   contains
     final     ::                b200Destructor
     procedure :: evolve      => b200Evolve
     procedure :: evolveBatch => b200EvolveBatch
  end type mergerTreeNodeEvolverB200

  interface mergerTreeNodeEvolverB200
     !!{RST
     Constructors for the :galacticus-class:`mergerTreeNodeEvolverB200` merger tree node evolver.
     !!}
     module procedure b200ConstructorParameters
  end interface mergerTreeNodeEvolverB200

contains

  function b200ConstructorParameters(parameters) result(self)
    !!{RST
    Constructor for the :galacticus-class:`mergerTreeNodeEvolverB200` merger tree node evolver class which takes a
    parameter set as input. All parameters of :galacticus-class:`mergerTreeNodeEvolverStandard` are accepted (and read by
    the parent constructor); the physics parameters are taken from the already-constructed objects of the parameter file.
    !!}
    use :: Input_Parameters, only : inputParameter, inputParameters
    use :: Error           , only : Error_Report
    implicit none
    type   (mergerTreeNodeEvolverB200)                :: self
    type   (inputParameters          ), intent(inout) :: parameters
    integer                                           :: status

    !![
    <inputParameter docformat="rst">
      <name>deviceOrdinal</name>
      <defaultValue>0</defaultValue>
      <variable>self%deviceOrdinal</variable>
      <description>The CUDA device on which this evolver runs (one evolver per host thread and :term:`GPU`).</description>
      <source>parameters</source>
    </inputParameter>
    <inputParameter docformat="rst">
      <name>resolveInterruptsOnDevice</name>
      <defaultValue>.true.</defaultValue>
      <variable>self%resolveInterruptsOnDevice</variable>
      <description>If true, component-creation interrupts (``hotHaloCreateByInterrupt``, ``diskCreateByInterrupt``, ``spheroidCreateByInterrupt``, ``blackHoleCreate``) are applied on the device and evolution of the node continues---equivalent to the interrupt loop of ``mergerTreeEvolverStandard``. Otherwise nodes are returned at the interrupt time.</description>
      <source>parameters</source>
    </inputParameter>
    !!]
    ! The parent reads odeToleranceAbsolute/Relative, odeAlgorithm, reuseODEStepSize, enforceNonNegativity, profileOdeEvolver.
    self%mergerTreeNodeEvolverStandard=mergerTreeNodeEvolverStandard(parameters)
    if (glc_abi_version() /= GLC_ABI_VERSION) call Error_Report('libglcb200 ABI version mismatch'//{introspection:location})
    status=glc_evolver_create(self%evolver,int(self%deviceOrdinal,c_int32_t))
    if (status /= 0) call Error_Report('no usable CUDA device: libglcb200 has no CPU fallback'//{introspection:location})
    call b200ParametersFill(self,parameters)
    status=glc_evolver_set_params(self%evolver,self%parameters_)
    if (status /= 0) call Error_Report('glc_evolver_set_params failed'//{introspection:location})
    call b200TablesUpload(self,parameters)
    !![
    <inputParametersValidate source="parameters"/>
    !!]
    return
  end function b200ConstructorParameters

  subroutine b200ParametersFill(self,parameters)
    !!{RST
    Fill the flat parameter structure of the C-ABI (``struct glc_params``: the field names are the :term:`XML` parameter
    names) from the objects built from the parameter file. Starts from ``glc_params_default`` (the values of
    ``parameters/quickTest.xml``) and overrides every field for which the parameter file has a value.
    !!}
    use :: Input_Parameters, only : inputParameters
    implicit none
    type   (mergerTreeNodeEvolverB200), intent(inout) :: self
    type   (inputParameters          ), intent(inout) :: parameters
    class  (cosmologyParametersClass ), pointer       :: cosmologyParameters_
    integer                                           :: status

    status=glc_params_default(self%parameters_,GLC_MODEL_STANDARD)
    ! mergerTreeNodeEvolverStandard (node_evolver/standard.F90:166-253)
    self%parameters_%odeToleranceAbsolute     =self%odeToleranceAbsolute
    self%parameters_%odeToleranceRelative     =self%odeToleranceRelative
    self%parameters_%reuseODEStepSize         =merge(1,0,self%reuseODEStepSize         )
    self%parameters_%enforceNonNegativity     =merge(1,0,self%enforceNonNegativity     )
    self%parameters_%profileOdeEvolver        =merge(1,0,self%profileOdeEvolver        )
    self%parameters_%resolveInterruptsOnDevice=merge(1,0,self%resolveInterruptsOnDevice)
    ! cosmologyParameters
    !![
    <objectBuilder class="cosmologyParameters" name="cosmologyParameters_" source="parameters"/>
    !!]
    self%parameters_%OmegaMatter   =cosmologyParameters_%OmegaMatter   (                  )
    self%parameters_%OmegaBaryon   =cosmologyParameters_%OmegaBaryon   (                  )
    self%parameters_%HubbleConstant=cosmologyParameters_%HubbleConstant(hubbleUnitsStandard)
    !![
    <objectDestructor name="cosmologyParameters_"/>
    !!]
    ! Physics classes: each value is read from the sub-parameters of the class the parameter file selects, e.g.
    !   [stellarPopulation/recycledFraction], [stellarPopulation/metalYield]
    !   [hotHaloMassDistribution/beta], [hotHaloMassDistributionCoreRadius/coreRadiusOverVirialRadius]
    !   [coolingRate/velocityCutOff], [coolingTime/degreesOfFreedom], [coolingInfallTorque/fractionLossAngularMomentum]
    !   [starFormationRateSurfaceDensityDisks/frequencyStarFormation, clumpingFactorMolecularComplex]
    !   [starFormationRateDisks/tolerance], [starFormationTimescale/efficiency, exponentVelocity, timescaleMinimum]
    !   [stellarFeedbackOutflows/.../velocityCharacteristic, exponent, timescaleOutflowFractionalMinimum] (disks, spheroids)
    !   [galacticStructureSolver/solutionTolerance, velocityMaximumFactor, includeBaryonGravity]
    !   [darkMatterProfile/A, omega], [darkMatterProfileDMO] (NFW | isothermal)
    !   [galacticDynamicsBarInstability/stabilityThresholdGaseous, stabilityThresholdStellar]
    !   [blackHoleSeeds/mass, spin], [blackHoleAccretionRate/*], [blackHoleWind/efficiencyWind, ...],
    !   [accretionDisks/accretionRateThinDiskMaximum, accretionRateThinDiskMinimum, accretionRateTransitionWidth, ...]
    !   [mergerTreeEvolver/timestepHostRelative, timestepHostAbsolute], [mergerTreeEvolveTimestep/timeStepRelative, timeStepAbsolute]
    ! with the same <inputParameter> blocks (names, defaults) those classes declare; the generated code of
    ! python/Galacticus/Build/SourceTree/Process/InputParameter.py does the look-up.
    call b200ParametersFillPhysics(self%parameters_,parameters)
    ! The node operator list: bit i of operatorMask enables operator i of enum glc_operator; an operator that is absent from
    ! <nodeOperator value="multi"> is masked, an operator with no device restatement is an error (no silent fallback).
    self%parameters_%operatorMask=b200OperatorMask(self%nodeOperator_)
    return
  end subroutine b200ParametersFill

  subroutine b200TablesUpload(self,parameters)
    !!{RST
    Hand the tabulated inputs to the library: the :term:`CIE` cooling function and electron fraction (the tables that
    ``cieFileReadFile`` holds, ``cooling/cooling_function/CIE_file.F90:535-663``), the mean virial density and its
    logarithmic growth rate on the lattice of ``virialDensityContrastDefinition`` (``virial_density_contrast.F90:356-417``),
    the exponential-disk rotation-curve factor (``exponential_disk.F90:675-733``) and the two :term:`ADAF` tabulations
    (``accretion_disks/ADAF.F90:394-447``).
    !!}
    use :: Input_Parameters, only : inputParameters
    use :: Error           , only : Error_Report
    implicit none
    type            (mergerTreeNodeEvolverB200), intent(inout)               :: self
    type            (inputParameters          ), intent(inout)               :: parameters
    double precision                           , allocatable, dimension(:  ), target :: x0, x1
    double precision                           , allocatable, dimension(:,:)         :: values
    integer                                                                  :: tableID, status

    do tableID=GLC_TABLE_COOLING_FUNCTION,GLC_NTABLES-1
       ! Each table is extracted from the object that owns it (coolingFunctionCIEFile%coolingFunctionTable,
       ! chemicalStateCIEFile%densityElectronTable, darkMatterHaloScaleVirialDensityContrastDefinition%meanDensityTable,
       ! massDistributionExponentialDisk%rotationCurveTable, accretionDisksADAF%tabulations) by b200TableExtract.
       call b200TableExtract(parameters,tableID,x0,x1,values)
       if (allocated(x1)) then
          status=glc_evolver_set_table(self%evolver,int(tableID,c_int32_t),int(size(x0),c_int32_t),int(size(x1),c_int32_t),x0,c_loc(x1),reshape(transpose(values),[size(values)]))
       else
          status=glc_evolver_set_table(self%evolver,int(tableID,c_int32_t),int(size(x0),c_int32_t),int(size(values,dim=2),c_int32_t),x0,c_null_ptr,reshape(transpose(values),[size(values)]))
       end if
       if (status /= 0) call Error_Report('glc_evolver_set_table failed'//{introspection:location})
    end do
    return
  end subroutine b200TablesUpload

  subroutine b200Destructor(self)
    !!{RST
    Destructor for the :galacticus-class:`mergerTreeNodeEvolverB200` class: releases the device arena and tables.
    !!}
    implicit none
    type   (mergerTreeNodeEvolverB200), intent(inout) :: self
    integer                                           :: status

    if (c_associated(self%evolver)) status=glc_evolver_destroy(self%evolver)
    self%evolver=c_null_ptr
    return
  end subroutine b200Destructor

  subroutine b200Evolve(self,tree,node,timeEnd,interrupted,functionInterrupt,galacticStructureSolver__,treeLock,systemClockMaximum,status)
    !!{RST
    Evolves ``node`` to time ``timeEnd``, or until evolution is interrupted: a batch of one node. Same interface and
    semantics as ``standardEvolve`` (``node_evolver/standard.F90:385``).
    !!}
    implicit none
    class           (mergerTreeNodeEvolverB200   ), intent(inout), target  :: self
    type            (mergerTree                  ), intent(inout)          :: tree
    type            (treeNode                    ), intent(inout), pointer :: node
    double precision                              , intent(in   )          :: timeEnd
    logical                                       , intent(  out)          :: interrupted
    procedure       (interruptTask               ), intent(  out), pointer :: functionInterrupt
    class           (galacticStructureSolverClass), intent(in   ), target  :: galacticStructureSolver__
    class           (ompLockClass                ), intent(inout)          :: treeLock
    integer         (kind_int8                   ), intent(in   ), optional:: systemClockMaximum
    integer                                       , intent(  out), optional:: status
    type            (treeNodeList                ), dimension(1)           :: nodes
    double precision                              , dimension(1)           :: timesEnd
    logical                                       , dimension(1)           :: interrupteds
    type            (interruptTaskList           ), dimension(1)           :: functionsInterrupt
    integer                                       , dimension(1)           :: statuses
    !$GLC attributes unused :: tree, galacticStructureSolver__, treeLock, systemClockMaximum

    nodes   (1)%node => node
    timesEnd(1)      =  timeEnd
    call self%evolveBatch(nodes,timesEnd,interrupteds,functionsInterrupt,statuses)
    interrupted       =  interrupteds      (1)
    functionInterrupt => functionsInterrupt(1)%task
    if (present(status)) status=statuses(1)
    return
  end subroutine b200Evolve

  subroutine b200EvolveBatch(self,nodes,timesEnd,interrupted,functionsInterrupt,status)
    !!{RST
    Evolve every node of ``nodes`` to its own end time in one call of ``glc_evolve_batch``.

    Gather: the numerically integrated properties go through ``node%serializeValues`` (the generated
    ``treeNodeSerializeValuesToArray``, ``python/Galacticus/Build/Components/TreeNodes/ODESolver.py:95-138``)---the record's
    first ``GLC_NY`` words *are* that array for the quickTest component set (checked against the generators by
    ``tests/test_layout.py``)---and the analytic / non-evolved words are read from the components and from the
    meta-properties of the interpolating node operators. Scatter is the inverse, followed by what ``standardEvolve`` does
    after its solve (``standard.F90:726-753``): the time, the step size guess, the interrupt procedure.
    !!}
    use :: Error           , only : Error_Report          , errorStatusSuccess
    use :: Galacticus_Nodes, only : nodeComponentBasic    , nodeComponentDisk      , nodeComponentSpheroid, nodeComponentHotHalo        , &
         &                          nodeComponentBlackHole, nodeComponentSatellite , nodeComponentSpin    , nodeComponentDarkMatterProfile, &
         &                          propertyTypeActive
    implicit none
    class           (mergerTreeNodeEvolverB200), intent(inout)               :: self
    type            (treeNodeList             ), intent(inout), dimension(:) :: nodes
    double precision                           , intent(in   ), dimension(:) :: timesEnd
    logical                                    , intent(  out), dimension(:) :: interrupted
    type            (interruptTaskList        ), intent(  out), dimension(:) :: functionsInterrupt
    integer                                    , intent(  out), dimension(:) :: status
    real            (c_double                 ), allocatable  , dimension(:,:) :: props
    integer         (c_int32_t                ), allocatable  , dimension(:  ) :: flags, statusDevice, interruptDevice
    type            (glcCounters              )                              :: counters
    integer                                                                  :: i, n, statusCall

    n=size(nodes)
    allocate(props(GLC_NPROP,n),flags(n),statusDevice(n),interruptDevice(n))
    do i=1,n
       call b200NodeGather(self,nodes(i)%node,props(:,i),flags(i))
    end do
    statusCall=glc_evolve_batch(self%evolver,int(n,c_int64_t),props,flags,timesEnd,statusDevice,interruptDevice,counters)
    if (statusCall /= 0) call Error_Report('glc_evolve_batch failed'//{introspection:location})
    do i=1,n
       call b200NodeScatter(self,nodes(i)%node,props(:,i),flags(i))
       ! glc_status values ARE the errorStatus* codes (source/error/_module.F90:66-75): no translation.
       status     (i)=int(statusDevice(i))
       ! A node that failed on its last trial comes back at its saved values: print what standardErrorHandler prints
       ! (node_evolver/standard.F90:1063-1140).
       if (statusDevice(i) == GLC_STATUS_UNDERFLOW) call b200ErrorReport(self,nodes(i)%node,props(:,i),flags(i))
       interrupted(i)=interruptDevice(i) /= GLC_INT_NONE
       select case (interruptDevice(i))
       case (GLC_INT_NONE           )
          functionsInterrupt(i)%task => null()
       case (GLC_INT_HOTHALO_CREATE )
          functionsInterrupt(i)%task => hotHaloCreateByInterrupt   ! python/Galacticus/Build/Components/Properties/Evolve.py:486-493
       case (GLC_INT_DISK_CREATE    )
          functionsInterrupt(i)%task => diskCreateByInterrupt
       case (GLC_INT_SPHEROID_CREATE)
          functionsInterrupt(i)%task => spheroidCreateByInterrupt
       case (GLC_INT_BH_CREATE      )
          functionsInterrupt(i)%task => blackHoleCreate            ! nodes/operators/physics/black_holes/seed.F90:179-180
       end select
    end do
    return
  end subroutine b200EvolveBatch

  subroutine b200ErrorReport(self,node,record,flags)
    !!{RST
    The "ODE system parameters" table of ``standardErrorHandler`` (``node_evolver/standard.F90:1063-1140``) for a node
    whose evolution failed, from ``glc_error_report_node``.
    !!}
    use :: Display           , only : displayIndent, displayMessage, displayUnindent
    use :: ISO_Varying_String, only : varying_string, assignment(=), operator(//)
    use :: Galacticus_Nodes  , only : propertyTypeActive
    implicit none
    class           (mergerTreeNodeEvolverB200), intent(inout)               :: self
    type            (treeNode                 ), intent(inout)               :: node
    real            (c_double                 ), intent(in   ), dimension(:) :: record
    integer         (c_int32_t                ), intent(in   )               :: flags
    type            (glcErrorReport           )                              :: report
    type            (varying_string           )                              :: message
    character       (len=12                   )                              :: label
    integer                                                                  :: i, j, statusCall

    statusCall=glc_error_report_node(self%evolver,record,flags,record(GLC_P_TIME_STEP+1),report)
    if (statusCall /= 0) return
    call node%serializeASCII()
    write (label,'(e12.6)') report%time
    message="time, timeStep = "//label
    write (label,'(e12.6)') report%time_step
    message=message//", "//label
    call displayMessage(message)
    call displayIndent('ODE system parameters')
    call displayMessage(' : y            : dy/dt        : yScale       : yError       : yErrorScaled')
    j=0
    do i=1,GLC_NY
       if (report%active(i) == 0) cycle
       j=j+1
       message=node%nameFromIndex(j,propertyTypeActive)
       write (label,'(e12.6)') report%y           (i)
       message=message//" : "//label
       write (label,'(e12.6)') report%dydt        (i)
       message=message//" : "//label
       write (label,'(e12.6)') report%scale       (i)
       message=message//" : "//label
       write (label,'(e12.6)') report%error       (i)
       message=message//" : "//label
       write (label,'(e12.6)') report%error_scaled(i)
       message=message//" : "//label
       call displayMessage(message)
    end do
    call displayUnindent('done')
    return
  end subroutine b200ErrorReport

  subroutine b200NodeGather(self,node,record,flags)
    !!{RST
    Serialize ``node`` into a node record. ``record`` is indexed from 1 here, the ``GLC_P_*`` enumerators from 0.
    !!}
    use :: Galacticus_Nodes, only : nodeComponentBasic, nodeComponentDisk     , nodeComponentSpheroid         , nodeComponentHotHalo  , &
         &                          nodeComponentSpin , nodeComponentBlackHole, nodeComponentDarkMatterProfile, nodeComponentSatellite, &
         &                          propertyTypeActive
    implicit none
    class           (mergerTreeNodeEvolverB200     ), intent(inout)               :: self
    type            (treeNode                      ), intent(inout)               :: node
    real            (c_double                      ), intent(  out), dimension(:) :: record
    integer         (c_int32_t                     ), intent(  out)               :: flags
    class           (nodeComponentBasic            ), pointer                     :: basic
    class           (nodeComponentDisk             ), pointer                     :: disk
    class           (nodeComponentSpheroid         ), pointer                     :: spheroid
    class           (nodeComponentHotHalo          ), pointer                     :: hotHalo
    class           (nodeComponentBlackHole        ), pointer                     :: blackHole
    class           (nodeComponentDarkMatterProfile), pointer                     :: darkMatterProfile
    class           (nodeComponentSpin             ), pointer                     :: spin
    type            (treeNode                      ), pointer                     :: nodeSatellite

    record=0.0d0
    flags =0
    basic             => node%basic            ()
    disk              => node%disk             ()
    spheroid          => node%spheroid         ()
    hotHalo           => node%hotHalo          ()
    blackHole         => node%blackHole        ()
    darkMatterProfile => node%darkMatterProfile()
    spin              => node%spin             ()
    ! Component existence (a component of the base class does not exist: cf. the "select type" tests throughout the operators).
    select type (hotHalo  )
    class is (nodeComponentHotHaloStandard  )
       flags=ior(flags,GLC_F_HAS_HOTHALO )
       if (hotHalo%isInitialized()) flags=ior(flags,GLC_F_HH_INITIALIZED)
    end select
    select type (disk     )
    class is (nodeComponentDiskStandard     )
       flags=ior(flags,GLC_F_HAS_DISK    )
    end select
    select type (spheroid )
    class is (nodeComponentSpheroidStandard )
       flags=ior(flags,GLC_F_HAS_SPHEROID)
    end select
    select type (blackHole)
    class is (nodeComponentBlackHoleStandard)
       flags=ior(flags,GLC_F_HAS_BH      )
    end select
    if (node%isSatellite()) flags=ior(flags,GLC_F_IS_SATELLITE)
    ! The numerically integrated properties, in the order of treeNodeSerializeValuesToArray for this component set. The
    ! offsets of components that do not exist are skipped by the generated code, so each block is written at its fixed place.
    if (iand(flags,GLC_F_HAS_BH      ) /= 0) then
       record(GLC_P_BH_MASS            +1)=blackHole%mass             ()
       record(GLC_P_BH_SPIN            +1)=blackHole%spin             ()
    end if
    if (iand(flags,GLC_F_HAS_DISK    ) /= 0) then
       record(GLC_P_DISK_MASS_STELLAR  +1)=disk     %massStellar      ()
       record(GLC_P_DISK_ABUND_STELLAR +1)=b200Metals(disk   %abundancesStellar     ())
       record(GLC_P_DISK_MASS_GAS      +1)=disk     %massGas          ()
       record(GLC_P_DISK_ABUND_GAS     +1)=b200Metals(disk   %abundancesGas         ())
       record(GLC_P_DISK_ANGMOM        +1)=disk     %angularMomentum  ()
       record(GLC_P_DISK_RADIUS        +1)=disk     %radius           ()
       record(GLC_P_DISK_VELOCITY      +1)=disk     %velocity         ()
    end if
    if (iand(flags,GLC_F_HAS_HOTHALO ) /= 0) then
       record(GLC_P_HH_MASS            +1)=hotHalo  %mass             ()
       record(GLC_P_HH_ABUND           +1)=b200Metals(hotHalo%abundances            ())
       record(GLC_P_HH_ANGMOM          +1)=hotHalo  %angularMomentum  ()
       record(GLC_P_HH_OUTFLOWED_MASS  +1)=hotHalo  %outflowedMass    ()
       record(GLC_P_HH_OUTFLOWED_ANGMOM+1)=hotHalo  %outflowedAngularMomentum()
       record(GLC_P_HH_OUTFLOWED_ABUND +1)=b200Metals(hotHalo%outflowedAbundances   ())
       record(GLC_P_HH_UNACCRETED_MASS +1)=hotHalo  %unaccretedMass   ()
       record(GLC_P_HH_UNACCRETED_ABUND+1)=b200Metals(hotHalo%unaccretedAbundances  ())
       record(GLC_P_HH_OUTER_RADIUS    +1)=hotHalo  %outerRadius      ()
       record(GLC_P_HH_STRIPPED_MASS   +1)=hotHalo  %strippedMass     ()
       record(GLC_P_HH_STRIPPED_ABUND  +1)=b200Metals(hotHalo%strippedAbundances    ())
    end if
    record   (GLC_P_SAT_BOUND_MASS     +1)=b200BoundMass(node)
    if (iand(flags,GLC_F_HAS_SPHEROID) /= 0) then
       record(GLC_P_SPH_MASS_STELLAR   +1)=spheroid %massStellar      ()
       record(GLC_P_SPH_ABUND_STELLAR  +1)=b200Metals(spheroid%abundancesStellar    ())
       record(GLC_P_SPH_MASS_GAS       +1)=spheroid %massGas          ()
       record(GLC_P_SPH_ABUND_GAS      +1)=b200Metals(spheroid%abundancesGas        ())
       record(GLC_P_SPH_ANGMOM         +1)=spheroid %angularMomentum  ()
       record(GLC_P_SPH_RADIUS         +1)=spheroid %radius           ()
       record(GLC_P_SPH_VELOCITY       +1)=spheroid %velocity         ()
    end if
    ! Analytic / non-evolved words. The targets and rates are the meta-properties of the interpolating operators:
    ! nodeOperatorDMOInterpolate (massDMOTarget, accretionRate: dark_matter_only_mass/interpolate.F90:216-238),
    ! nodeOperatorDarkMatterProfileScaleInterpolate, nodeOperatorHaloAngularMomentumInterpolate.
    record(GLC_P_TIME              +1)=basic            %time            ()
    record(GLC_P_TIME_STEP         +1)=node             %timeStep        ()
    record(GLC_P_BASIC_MASS        +1)=basic            %mass            ()
    record(GLC_P_MASS_RATE         +1)=basic            %accretionRate   ()
    record(GLC_P_TIME_LAST_ISOLATED+1)=basic            %timeLastIsolated()
    record(GLC_P_DMSCALE           +1)=darkMatterProfile%scale           ()
    record(GLC_P_DMSCALE_RATE      +1)=darkMatterProfile%scaleGrowthRate ()
    record(GLC_P_SPIN              +1)=spin             %angularMomentum ()
    record(GLC_P_SPIN_RATE         +1)=spin             %angularMomentumGrowthRate()
    call b200InterpolationTargets(node,record(GLC_P_MASS_TARGET+1),record(GLC_P_TIME_TARGET+1),record(GLC_P_DMSCALE_TARGET+1),record(GLC_P_SPIN_TARGET+1))
    ! Baryonic mass of all sub-satellites, frozen for the call (dark_matter_profiles/adiabatic_Gnedin2004.F90:327-347).
    nodeSatellite => node%firstSatellite
    do while (associated(nodeSatellite))
       record(GLC_P_MASS_BARYONIC_SUBHALOS+1)=record(GLC_P_MASS_BARYONIC_SUBHALOS+1)+nodeSatellite%massBaryonic()
       nodeSatellite => nodeSatellite%sibling
    end do
    return
  end subroutine b200NodeGather

  subroutine b200NodeScatter(self,node,record,flags)
    !!{RST
    Deserialize a node record into ``node``: the inverse of ``b200NodeGather`` for everything the device changes, including
    components created on the device (``resolveInterruptsOnDevice``): these are created here with the generated
    ``<class>CreateByInterrupt`` procedures before their values are set.
    !!}
    use :: Galacticus_Nodes, only : nodeComponentBasic, nodeComponentDisk, nodeComponentSpheroid, nodeComponentHotHalo, nodeComponentBlackHole
    implicit none
    class           (mergerTreeNodeEvolverB200), intent(inout)               :: self
    type            (treeNode                 ), intent(inout), target       :: node
    real            (c_double                 ), intent(in   ), dimension(:) :: record
    integer         (c_int32_t                ), intent(in   )               :: flags
    class           (nodeComponentBasic       ), pointer                     :: basic
    class           (nodeComponentDisk        ), pointer                     :: disk
    class           (nodeComponentSpheroid    ), pointer                     :: spheroid
    class           (nodeComponentHotHalo     ), pointer                     :: hotHalo
    class           (nodeComponentBlackHole   ), pointer                     :: blackHole

    basic => node%basic()
    if (iand(flags,GLC_F_HAS_HOTHALO ) /= 0) then
       hotHalo   => node%hotHalo  (autoCreate=.true.)
       call hotHalo  %                   massSet(            record(GLC_P_HH_MASS            +1) )
       call hotHalo  %             abundancesSet(b200Abundances(record(GLC_P_HH_ABUND         +1)))
       call hotHalo  %        angularMomentumSet(            record(GLC_P_HH_ANGMOM          +1) )
       call hotHalo  %          outflowedMassSet(            record(GLC_P_HH_OUTFLOWED_MASS  +1) )
       call hotHalo  %outflowedAngularMomentumSet(           record(GLC_P_HH_OUTFLOWED_ANGMOM+1) )
       call hotHalo  %    outflowedAbundancesSet(b200Abundances(record(GLC_P_HH_OUTFLOWED_ABUND+1)))
       call hotHalo  %         unaccretedMassSet(            record(GLC_P_HH_UNACCRETED_MASS +1) )
       call hotHalo  %   unaccretedAbundancesSet(b200Abundances(record(GLC_P_HH_UNACCRETED_ABUND+1)))
       call hotHalo  %            outerRadiusSet(            record(GLC_P_HH_OUTER_RADIUS    +1) )
       call hotHalo  %           strippedMassSet(            record(GLC_P_HH_STRIPPED_MASS   +1) )
       call hotHalo  %     strippedAbundancesSet(b200Abundances(record(GLC_P_HH_STRIPPED_ABUND+1)))
       if (iand(flags,GLC_F_HH_INITIALIZED) /= 0) call hotHalo%isInitializedSet(.true.)
    end if
    if (iand(flags,GLC_F_HAS_DISK    ) /= 0) then
       disk      => node%disk     (autoCreate=.true.)
       call disk     %            massStellarSet(            record(GLC_P_DISK_MASS_STELLAR  +1) )
       call disk     %      abundancesStellarSet(b200Abundances(record(GLC_P_DISK_ABUND_STELLAR+1)))
       call disk     %                massGasSet(            record(GLC_P_DISK_MASS_GAS      +1) )
       call disk     %          abundancesGasSet(b200Abundances(record(GLC_P_DISK_ABUND_GAS   +1)))
       call disk     %        angularMomentumSet(            record(GLC_P_DISK_ANGMOM        +1) )
       call disk     %                 radiusSet(            record(GLC_P_DISK_RADIUS        +1) )
       call disk     %               velocitySet(            record(GLC_P_DISK_VELOCITY      +1) )
    end if
    if (iand(flags,GLC_F_HAS_SPHEROID) /= 0) then
       spheroid  => node%spheroid (autoCreate=.true.)
       call spheroid %            massStellarSet(            record(GLC_P_SPH_MASS_STELLAR   +1) )
       call spheroid %      abundancesStellarSet(b200Abundances(record(GLC_P_SPH_ABUND_STELLAR+1)))
       call spheroid %                massGasSet(            record(GLC_P_SPH_MASS_GAS       +1) )
       call spheroid %          abundancesGasSet(b200Abundances(record(GLC_P_SPH_ABUND_GAS    +1)))
       call spheroid %        angularMomentumSet(            record(GLC_P_SPH_ANGMOM         +1) )
       call spheroid %                 radiusSet(            record(GLC_P_SPH_RADIUS         +1) )
       call spheroid %               velocitySet(            record(GLC_P_SPH_VELOCITY       +1) )
    end if
    if (iand(flags,GLC_F_HAS_BH      ) /= 0) then
       blackHole => node%blackHole(autoCreate=.true.)
       call blackHole%                   massSet(            record(GLC_P_BH_MASS            +1) )
       call blackHole%                   spinSet(            record(GLC_P_BH_SPIN            +1) )
    end if
    call b200BoundMassSet(node,record(GLC_P_SAT_BOUND_MASS+1))
    ! standard.F90:726-740: analytic properties at the time reached, the time itself, the step size guess.
    call self %nodeOperator_%differentialEvolutionSolveAnalytics(node,record(GLC_P_TIME+1))
    call basic%timeSet    (record(GLC_P_TIME     +1))
    call node %timeStepSet(record(GLC_P_TIME_STEP+1))
    return
  end subroutine b200NodeScatter
